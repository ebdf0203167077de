"""GPU tests of the device-side synthetic generator (BASELINE config C5) and of the hypersparse
regime it produces: row-block shards of a power-law matrix with far more columns than one x tile
holds, generated and formatted on the device, multiplied through the C ABI and compared with the
oracle on the host. Bit-exact for fixed point; fp32 within 1e-5 * sum|a_i x_i| of fp64."""
import numpy as np
import pytest

from hisparse_b200 import capi, matgen

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def gpu():
    if capi.device_count() < 1:
        pytest.fail("no CUDA device: the GPU tests must run on the B200 box (no CPU fallback exists)")
    return True


def _csr_ok(ip, ix, rows, cols):
    assert ip[0] == 0 and ip.size == rows + 1 and ip[-1] == ix.size
    d = np.diff(ip.astype(np.int64))
    assert np.all(d >= 0)
    assert ix.size == 0 or int(ix.max()) < cols
    # sorted and unique inside every row: a descent may only happen at a row boundary
    desc = np.nonzero(np.diff(ix.astype(np.int64)) <= 0)[0] + 1
    assert np.all(np.isin(desc, ip)), "column ids not strictly increasing inside a row"
    return d


def test_generator_shape_and_determinism(gpu):
    rows, cols = 40000, 3_000_000
    m = capi.DeviceCsr.powerlaw(0, rows, cols, mean_degree=20.0, max_degree=5000, band_half_width=1 << 16, seed=11)
    ip, ix, vv = m.download()
    deg = _csr_ok(ip, ix, rows, cols)
    assert 12.0 < deg.mean() < 30.0 and deg.max() <= 5000
    # 80 % of the draws fall inside the band around the diagonal
    r = np.repeat(np.arange(rows, dtype=np.int64), deg)
    dist = np.abs(ix.astype(np.int64) - r)
    dist = np.minimum(dist, cols - dist)
    assert 0.7 < float((dist <= (1 << 16)).mean()) < 0.9
    # values are U[0,1) fp32
    v = vv.view(np.float32)
    assert v.min() >= 0.0 and v.max() < 1.0 and 0.4 < float(v.mean()) < 0.6
    # a shard is a pure function of (seed, global row): two half shards == the whole
    a = capi.DeviceCsr.powerlaw(0, 16000, cols, first_global_row=0, mean_degree=20.0, max_degree=5000,
                                band_half_width=1 << 16, seed=11)
    b = capi.DeviceCsr.powerlaw(0, rows - 16000, cols, first_global_row=16000, mean_degree=20.0, max_degree=5000,
                                band_half_width=1 << 16, seed=11)
    ipa, ixa, va = a.download()
    ipb, ixb, vb = b.download()
    assert np.array_equal(np.concatenate([ixa, ixb]), ix) and np.array_equal(np.concatenate([va, vb]), vv)
    assert np.array_equal(np.concatenate([ipa[:-1], ipb + ipa[-1]]), ip)
    for h in (m, a, b):
        h.free()


@pytest.mark.parametrize("impl,rpp", [("fixed", 0), ("fixed", 8192), ("float_pob", 0)])
def test_hypersparse_shard_parity(gpu, port, impl, rpp):
    """rows x 5 M columns, ~20 non-zeros per row: ~90 column tiles, nearly every (row, tile) segment
    is a single non-zero -- the regime of the C5 shards, at a size the oracle finishes at once."""
    rows, cols = 30000 + 64, 5_000_000
    fixed = impl == "fixed"
    m = capi.DeviceCsr.powerlaw(0, rows, cols, first_global_row=123456, mean_degree=20.0, max_degree=100000,
                                band_half_width=1 << 18, seed=5, q824=fixed, value_scale=0.05 if fixed else 1.0)
    ip, ix, vv = m.download()
    _csr_ok(ip, ix, rows, cols)
    ctx = capi.Context(0, impl)
    ctx.upload_matrix_csr_device(m, rpp)
    xf = np.random.default_rng(3).random(cols, dtype=np.float32)
    xw = matgen.quantize_q824(xf) if fixed else xf.view(np.uint32)
    ctx.upload_vector(xw)
    ctx.spmv()
    y = ctx.download_result()
    st = ctx.stats()
    assert st["nnz"] == m.nnz
    if fixed:
        assert np.array_equal(y, port.spmv_q824(ip, ix, vv, xw))
        # second SpMV with another vector on the same resident matrix (x = 1.0: products are the values)
        ones = np.full(cols, 1 << 24, np.uint32)
        ctx.upload_vector(ones)
        ctx.spmv()
        y1 = ctx.download_result()
        sums = np.add.reduceat(np.concatenate([vv.astype(np.uint64), [0]]), np.minimum(ip[:-1], vv.size))
        sums[np.diff(ip.astype(np.int64)) == 0] = 0
        assert np.array_equal(y1, np.minimum(sums, 0xFFFFFFFF).astype(np.uint32))
    else:
        y64, sa = port.spmv_f64(ip, ix, vv.view(np.float32), xf)
        assert np.all(np.abs(y.view(np.float32).astype(np.float64) - y64) <= TOL * sa + 1e-30)
    ctx.close()
    m.free()


@pytest.mark.parametrize("impl", ["fixed", "float_pob"])
def test_hypersparse_million_row_shard(gpu, port, impl):
    """A 2^20-row block of the C5 matrix itself (100 M columns, 1744 column tiles of 57,344, ~20 M non-zeros, generated and
    formatted on the device): every row against the oracle -- fixed point bit for bit, fp32 within 1e-5 norm-wise."""
    rows, cols = 1 << 20, 100_000_000
    fixed = impl == "fixed"
    m = capi.DeviceCsr.powerlaw(0, rows, cols, first_global_row=37_500_000, seed=0xC0FFEE05, q824=fixed,
                                value_scale=0.05 if fixed else 1.0)
    ip, ix, vv = m.download()
    # a window fetched on its own is the same piece of the matrix
    wip, wix, wv = m.download_rows(1000, 5000)
    assert np.array_equal(wip, ip[1000:5001] - ip[1000]) and np.array_equal(wix, ix[ip[1000]:ip[5000]])
    assert np.array_equal(wv, vv[ip[1000]:ip[5000]])
    ctx = capi.Context(0, impl)
    ctx.upload_matrix_csr_device(m)
    m.free()
    xf = np.random.default_rng(9).random(cols, dtype=np.float32)
    xw = matgen.quantize_q824(xf) if fixed else xf.view(np.uint32)
    ctx.upload_vector(xw)
    ctx.spmv()
    y = ctx.download_result()
    assert ctx.stats()["n_col_tiles"] >= 1744 and ctx.stats()["layout"] == 1
    if fixed:
        assert np.array_equal(y, port.spmv_q824(ip, ix, vv, xw))
    else:
        y64, sa = port.spmv_f64(ip, ix, vv.view(np.float32), xf)
        assert np.all(np.abs(y.view(np.float32).astype(np.float64) - y64) <= TOL * sa + 1e-30)
    ctx.close()
