"""The C restatement (oracle/hsoracle.c) against the reference's own code (oracle/_ref):
Q8.24 arithmetic, CSR->CPSR formatter, per-channel images, the dataflow top_wrapper and
compute_ref. CPU only. These tests are what "pins" the oracle."""
import numpy as np
import pytest

from hisparse_b200 import matgen
from oracle import hsoracle

IMPL_ID = {"fixed": 0, "float_pob": 1, "float_stall": 2}


def _values_for(impl, data, port):
    """32-bit value words the kernel sees: Q8.24 raw for fixed, IEEE bits for float."""
    if impl == "fixed":
        return port.quantize(data)
    return np.ascontiguousarray(data, np.float32).view(np.uint32)


def _mats():
    out = []
    out.append(("dense128", matgen.dense_csr(128, 128), False))
    out.append(("uniform1000", matgen.uniform_sparse_csr(1000, 1024, 10), False))
    r = matgen.random_csr(700, 40000, 0.002, 11)          # 2 column partitions, ragged rows
    out.append(("rand700x40000", r, False))
    r = matgen.random_csr(3000, 500, 0.004, 12)           # many empty rows
    out.append(("sparse_skip", r, True))
    r = matgen.rmat_csr(5000, 60000, 13)                   # power law, empty + long rows
    out.append(("rmat5000_skip", r, True))
    out.append(("rmat5000_noskip", r, False))
    return out


MATS = _mats()


def test_quantize_matches_reference(port, refs):
    rng = np.random.default_rng(0)
    v = np.concatenate([rng.random(20000, dtype=np.float32) * 300 - 10,
                        np.array([0, 1, 255.99999, 256, 1e9, -1, 2 ** -25, 3 * 2 ** -25, 2 ** -24, 0.5 + 2 ** -25],
                                 np.float32),
                        (rng.integers(0, 2 ** 31, 5000) * 2.0 ** -24).astype(np.float32) + np.float32(2 ** -25)])
    assert np.array_equal(port.quantize(v), refs["fixed"].val_from_float(v))


def test_mac_chain_matches_reference(port, refs):
    rng = np.random.default_rng(1)
    for scale in (2 ** 8, 2 ** 20, 2 ** 26, 2 ** 32):
        for n in (1, 7, 300):
            a = rng.integers(0, scale, n, dtype=np.uint64).astype(np.uint32)
            b = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
            assert port.mac_chain(a, b) == refs["fixed"].mac_chain(a, b)
    # rounding boundary: product fraction exactly one half, and saturation of product and of sum
    a = np.array([0x00800000, 0xFFFFFFFF, 0xFFFFFFFF, 1], np.uint32)
    b = np.array([0x00000001, 0xFFFFFFFF, 0x01000000, 0x00800000], np.uint32)
    for k in range(1, 5):
        assert port.mac_chain(a[:k], b[:k]) == refs["fixed"].mac_chain(a[:k], b[:k])


@pytest.mark.parametrize("impl", hsoracle.IMPLS)
@pytest.mark.parametrize("name,mat,skip", MATS, ids=[m[0] for m in MATS])
def test_formatter_and_dataflow(port, refs, impl, name, mat, skip):
    ref = refs[impl]
    rows, cols, indptr, indices, data = mat
    IF, P, NCH = ref.INTERLEAVE_FACTOR, ref.PACK_SIZE, ref.NUM_HBM_CHANNELS
    # small buffers so that several row/column partitions occur on small matrices
    ob = 2 * P * NCH * IF
    vb = 256 if cols <= 2048 else ref.LOGICAL_VB_SIZE
    if name.startswith("rmat"):
        ob, vb = 4 * P * NCH * IF, 2048
    r2, c2, ip2 = matgen.pad_csr(rows, cols, indptr, P * NCH * IF, P)
    words = _values_for(impl, data, port)
    kind = hsoracle.VAL_Q824 if impl == "fixed" else hsoracle.VAL_FLOAT_BITS

    # 1. formatter: every (row part, col part, virtual channel) block identical to the reference's
    mine = port.csr2cpsr(r2, c2, ip2, indices, words, P, ob, vb, NCH * IF, skip, kind)
    theirs = ref.csr2cpsr(rows, cols, indptr, indices, data, skip, ob=ob, vb=vb)
    assert (mine.rows, mine.cols, mine.n_row_parts, mine.n_col_parts) == \
           (theirs.rows, theirs.cols, theirs.n_row_parts, theirs.n_col_parts)
    for j in range(mine.n_row_parts):
        for i in range(mine.n_col_parts):
            for c in range(NCH * IF):
                a_idx, a_val, a_ptr = mine.block(j, i, c)
                b_idx, b_val, b_len = theirs.block(j, i, c)
                assert np.array_equal(a_idx, b_idx), (j, i, c)
                assert np.array_equal(a_val, b_val), (j, i, c)
                assert np.array_equal(a_ptr[-1], b_len), (j, i, c)

    # The kernels have the partition sizes compiled in (vector loader / result drain use
    # LOGICAL_VB_SIZE / LOGICAL_OB_SIZE, spmv_vector_loader.cpp:13-19, spmv_result_drain.cpp:36),
    # so the dataflow run uses the hardware sizes.
    OB, VB = ref.LOGICAL_OB_SIZE, ref.LOGICAL_VB_SIZE
    full = port.csr2cpsr(r2, c2, ip2, indices, words, P, OB, VB, NCH * IF, skip, kind)
    images = full.channel_images(IF)
    ncp, nrp = full.n_col_parts, full.n_row_parts
    rng = np.random.default_rng(5)
    xf = rng.random(c2, dtype=np.float32) if impl == "fixed" else (rng.random(c2, dtype=np.float32) * 2 - 1)
    xw = _values_for(impl, xf, port)
    y_ref = np.zeros(r2, np.uint32)
    y_port = np.zeros(r2, np.uint32)
    for rp in range(nrp):
        rows_here = OB if rp < nrp - 1 or r2 % OB == 0 else r2 % OB
        part_len = rows_here // NCH
        ref.top_wrapper(images, xw, y_ref, rp, part_len, ncp, ncp * nrp, c2)
        port.top_wrapper(images, xw, y_port, IMPL_ID[impl], IF, OB, VB, rp, part_len, ncp, ncp * nrp, c2)

    if impl == "fixed":
        # 2. bit-exact: reference dataflow == functional restatement == closed form on CSR
        y_csr = port.spmv_q824(ip2, indices, words, xw)
        assert np.array_equal(y_ref, y_port)
        assert np.array_equal(y_ref, y_csr)
    else:
        # float: accumulation order inside the reference is a property of its shuffle timing;
        # compare within tolerance of the fp64 result (norm-wise 1e-5)
        y64, sa = port.spmv_f64(ip2, indices, data, xf)
        for y in (y_ref, y_port):
            err = np.abs(y.view(np.float32).astype(np.float64) - y64)
            assert np.all(err <= 1e-5 * sa + 1e-30)


def test_compute_ref_matches_reference(port, refs):
    for name, mat, _ in MATS:
        rows, cols, indptr, indices, data = mat
        rng = np.random.default_rng(3)
        d = (rng.random(data.size, dtype=np.float32) - 0.3).astype(np.float32)
        x = rng.random(cols, dtype=np.float32)
        a = port.spmv_f32(indptr, indices, d, x)
        b = refs["fixed"].compute_ref(rows, cols, indptr, indices, d, x)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), name


def test_reference_harness_passes_under_shim(refs):
    """The reference's own end-to-end harness (its layout code, rand() x, verify 1e-4) accepts
    the shim for all three implementations, with and without skip-empty-rows."""
    rows, cols, indptr, indices, data = matgen.rmat_csr(3000, 30000, 21, values="ones")
    data = data * np.float32(0.01)
    for impl in hsoracle.IMPLS:
        assert refs[impl].test_harness(rows, cols, indptr, indices, data, False)
        assert refs[impl].test_harness(rows, cols, indptr, indices, data, True)


def test_reference_selftests(refs):
    for impl in hsoracle.IMPLS:
        assert refs[impl].selftest()
