"""Synthetic CSR generators for the SpMV workloads of BASELINE.json.

The reference's datasets (datasets/graph, datasets/pruned_nn) are download targets that are not
shipped (datasets/download.sh); these generators produce stand-ins of the same shape, nnz and
degree distribution. All generators are seeded and return (rows, cols, indptr, indices, data)
with uint32 indices, float32 data, column ids sorted and unique inside every row.

`dense_csr` and `uniform_sparse_csr` restate the reference's own synthetic test matrices
(sw/host.cpp:382-430 == spmv_csim/csim.cpp:387-435).
"""
import numpy as np


def _finish(rows, cols, r, c, rng, values):
    key = r.astype(np.uint64) * np.uint64(cols) + c.astype(np.uint64)
    # sorted distinct keys; (numpy 2.3's np.unique takes a minute for 6e7 64-bit keys, sort + mask a second)
    key.sort()
    if key.size:
        keep = np.empty(key.size, bool)
        keep[0] = True
        np.not_equal(key[1:], key[:-1], out=keep[1:])
        key = key[keep]
    r = (key // np.uint64(cols)).astype(np.uint32)
    c = (key % np.uint64(cols)).astype(np.uint32)
    indptr = np.zeros(rows + 1, np.uint32)
    np.cumsum(np.bincount(r, minlength=rows), out=indptr[1:])
    nnz = c.size
    if values == "ones":
        data = np.ones(nnz, np.float32)
    elif values == "u01":
        data = rng.random(nnz, dtype=np.float32)
    elif values == "normal":
        data = (0.05 * rng.standard_normal(nnz)).astype(np.float32)
    elif values == "small":
        data = (rng.random(nnz, dtype=np.float32) * np.float32(0.01)).astype(np.float32)
    else:
        raise ValueError(values)
    return rows, cols, indptr, c, data


def dense_csr(rows, cols):
    """create_dense_CSR (sw/host.cpp:382-405): all ones."""
    indptr = (np.arange(rows + 1, dtype=np.uint64) * cols).astype(np.uint32)
    indices = np.tile(np.arange(cols, dtype=np.uint32), rows)
    return rows, cols, indptr, indices, np.ones(rows * cols, np.float32)


def uniform_sparse_csr(rows, cols, nnz_per_row):
    """create_uniform_sparse_CSR (sw/host.cpp:407-430): col = (step*j + i) % cols, all ones.
    NOTE: like the reference, column ids inside a row are NOT sorted when they wrap."""
    step = cols // nnz_per_row
    i = np.arange(rows, dtype=np.uint64)[:, None]
    j = np.arange(nnz_per_row, dtype=np.uint64)[None, :]
    indices = ((step * j + i) % cols).astype(np.uint32).ravel()
    indptr = (np.arange(rows + 1, dtype=np.uint64) * nnz_per_row).astype(np.uint32)
    return rows, cols, indptr, indices, np.ones(rows * nnz_per_row, np.float32)


def random_csr(rows, cols, density, seed, values="u01"):
    """Uniform random sparsity (config C1: 4096 x 4096, 1 %)."""
    rng = np.random.default_rng(seed)
    n = int(round(rows * cols * density))
    m = int(n * 1.02) + 16
    r = rng.integers(0, rows, m, dtype=np.uint32)
    c = rng.integers(0, cols, m, dtype=np.uint32)
    return _finish(rows, cols, r, c, rng, values)


def rmat_csr(n, nnz_target, seed, a=0.57, b=0.19, c=0.19, values="u01", symmetric=False,
             oversample=1.36):
    """R-MAT power-law graph (stand-in for googleplus / ogbl-ppa, configs C2 / C4)."""
    rng = np.random.default_rng(seed)
    levels = int(np.ceil(np.log2(n)))
    m = int(nnz_target * oversample) + 1024
    if symmetric:
        m //= 2
    # 16-bit resolution categorical draw per level through lookup tables
    t = np.arange(65536, dtype=np.float64) / 65536.0
    lut_r = (t >= a + b).astype(np.uint32)
    lut_c = (((t >= a) & (t < a + b)) | (t >= a + b + c)).astype(np.uint32)
    r = np.zeros(m, np.uint32)
    col = np.zeros(m, np.uint32)
    for _ in range(levels):
        u = rng.integers(0, 65536, m, dtype=np.uint16)
        r <<= np.uint32(1)
        r |= lut_r[u]
        col <<= np.uint32(1)
        col |= lut_c[u]
    ok = (r < n) & (col < n)
    r, col = r[ok], col[ok]
    # scatter the hubs: R-MAT concentrates mass at low ids; a fixed permutation keeps the
    # degree distribution while removing the artificial locality
    perm = rng.permutation(n).astype(np.uint32)
    r, col = perm[r], perm[col]
    if symmetric:
        r, col = np.concatenate([r, col]), np.concatenate([col, r])
    out = _finish(n, n, r, col, rng, values)
    return out


def bernoulli_csr(rows, cols, density, seed, values="normal"):
    """Unstructured-pruning mask (config C3: transformer 512 x 33288, density 5..50 %)."""
    rng = np.random.default_rng(seed)
    mask = rng.random((rows, cols), dtype=np.float32) < density
    r, c = np.nonzero(mask)
    return _finish(rows, cols, r.astype(np.uint32), c.astype(np.uint32), rng, values)


def pad_csr(rows, cols, indptr, row_div, col_div):
    """util_round_csr_matrix_dim (sw/data_formatter.h:15-29) on the index arrays."""
    r2 = rows + (-rows) % row_div
    c2 = cols + (-cols) % col_div
    ip = np.concatenate([indptr, np.full(r2 - rows, indptr[-1], np.uint32)]).astype(np.uint32)
    return r2, c2, ip


def quantize_q824(v):
    """float -> raw Q8.24 words the way the host conversion does it (spmv::ufixed_q8_24 in
    hisparse_b200/host/fixed_point.h == csr_matrix_convert_from_float<VAL_T>, sw/data_loader.h:76-84):
    round half up at 2^-24, clamp to [0, 2^32 - 1]; negatives and NaN give 0."""
    d = np.asarray(v, dtype=np.float64)
    s = np.floor(np.ldexp(d, 24) + 0.5)
    s = np.where(d > 0, s, 0.0)
    return np.minimum(s, 4294967295.0).astype(np.uint32)
