"""GPU parity tests: the sm_100a path, called through the C ABI (include/hisparse_b200.h), against
  * the oracle restatement (oracle/hsoracle.c) on seeded inputs,
  * the committed fixtures the reference C simulation produced (tests/golden/), and
  * the reference's own code (oracle/_ref) when the prebuilt library is present.
Bit-exact for the fixed-point path; fp32 within 1e-5 norm-wise of the reference CPU SpMV
(|y - y_ref| <= 1e-5 * sum_i |a_i x_i|, the tolerance BASELINE.json states, made norm-wise so
that it is meaningful for rows with cancellation)."""
import glob
import os

import numpy as np
import pytest

from hisparse_b200 import capi, matgen
from oracle import hsoracle

pytestmark = pytest.mark.gpu

FIXTURES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "csim_*.npz")))
TOL = 1e-5


@pytest.fixture(scope="module")
def gpu():
    if capi.device_count() < 1:
        pytest.fail("no CUDA device: the GPU tests must run on the B200 box (no CPU fallback exists)")
    return True


def run_fixed_csr(port, mat, x_f32, rows_per_partition=0):
    rows, cols, indptr, indices, data = mat
    words = port.quantize(data)
    xw = port.quantize(x_f32)
    ctx = capi.Context(0, capi.IMPL_FIXED)
    ctx.upload_matrix_csr(rows, cols, indptr, indices, words, rows_per_partition)
    ctx.upload_vector(xw)
    ctx.spmv()
    y = ctx.download_result()
    st = ctx.stats()
    ctx.close()
    return y, port.spmv_q824(indptr, indices, words, xw), st


def check_float(y_words, port, indptr, indices, data, x, extra_ref=None):
    y = y_words.view(np.float32).astype(np.float64)
    y64, sa = port.spmv_f64(indptr, indices, data, x)
    yref = port.spmv_f32(indptr, indices, data, x).astype(np.float64)   # compute_ref restatement
    bound = TOL * sa + 1e-30
    assert np.all(np.abs(y - yref) <= bound), float(np.max(np.abs(y - yref) / (sa + 1e-30)))
    assert np.all(np.abs(y - y64) <= bound)


# ------------------------------------------------------------------------------------------
# fixed point, CSR fast path
# ------------------------------------------------------------------------------------------
FIXED_CASES = [
    ("dense128", lambda: matgen.dense_csr(128, 128), 0, 1.0),
    ("uniform_1000x1024", lambda: matgen.uniform_sparse_csr(1000, 1024, 10), 0, 1.0),
    ("rand_4096_1pct", lambda: matgen.random_csr(4096, 4096, 0.01, 0xC0FFEE01), 0, 1.0),
    ("multi_tile_70000", lambda: matgen.random_csr(900, 70000, 0.003, 7), 0, 1.0),
    ("rmat_20000", lambda: matgen.rmat_csr(20000, 600000, 8), 0, 1.0),
    ("rmat_row_partitions", lambda: matgen.rmat_csr(20000, 300000, 9), 4096, 1.0),
    ("saturating", lambda: matgen.random_csr(512, 3000, 0.2, 10), 0, 60.0),
    ("more_chunks_than_sms", lambda: matgen.random_csr(3000, 3000, 0.03, 11), 0, 1.0),
]


@pytest.mark.parametrize("name,make,rpp,scale", FIXED_CASES, ids=[c[0] for c in FIXED_CASES])
def test_fixed_csr_bit_exact(gpu, port, name, make, rpp, scale):
    rows, cols, indptr, indices, data = make()
    rng = np.random.default_rng(1)
    data = (data * np.float32(scale)).astype(np.float32)
    x = (rng.random(cols, dtype=np.float32) * np.float32(scale)).astype(np.float32)
    y, want, st = run_fixed_csr(port, (rows, cols, indptr, indices, data), x, rpp)
    assert np.array_equal(y, want)
    if name == "saturating":
        assert (want == 0xFFFFFFFF).any()
    if rpp:
        assert st["n_row_parts"] == (rows + rpp - 1) // rpp


def test_fixed_edge_cases(gpu, port):
    # empty matrix, all-empty rows, a single very long row crossing two tiles, x of zeros
    ip = np.zeros(129, np.uint32)
    y, want, _ = run_fixed_csr(port, (128, 64, ip, np.zeros(0, np.uint32), np.zeros(0, np.float32)),
                               np.ones(64, np.float32))
    assert np.array_equal(y, want) and not y.any()
    n = 70000
    ip = np.array([0, 0, n, n, n + 1], np.uint32)
    ix = np.concatenate([np.arange(n), [3]]).astype(np.uint32)
    d = np.full(n + 1, 0.001, np.float32)
    y, want, st = run_fixed_csr(port, (4, n, ip, ix, d), np.full(n, 0.5, np.float32))
    assert np.array_equal(y, want) and st["n_col_tiles"] == 2     # 70000 columns, dense rows: two tiles of <= 44000
    y, want, _ = run_fixed_csr(port, (4, n, ip, ix, d), np.zeros(n, np.float32))
    assert np.array_equal(y, want) and not y.any()
    # reference-style x in {0,1} (sw/host.cpp:238)
    rows, cols, indptr, indices, data = matgen.rmat_csr(5000, 80000, 3)
    x = (np.random.default_rng(0).integers(0, 2, cols)).astype(np.float32)
    y, want, _ = run_fixed_csr(port, (rows, cols, indptr, indices, data), x)
    assert np.array_equal(y, want)


def test_repeated_runs_and_vector_update(gpu, port):
    rows, cols, indptr, indices, data = matgen.rmat_csr(8000, 200000, 21)
    words = port.quantize(data)
    ctx = capi.Context(0, capi.IMPL_FIXED)
    ctx.upload_matrix_csr(rows, cols, indptr, indices, words)
    ctx.set_replicas(3)
    rng = np.random.default_rng(4)
    for it in range(5):
        xw = port.quantize(rng.random(cols, dtype=np.float32))
        ctx.upload_vector(xw)
        ctx.spmv()
        assert np.array_equal(ctx.download_result(), port.spmv_q824(indptr, indices, words, xw)), it
    assert ctx.stats()["kernel_launches"] == 10      # per SpMV: one fused launch + the drain forced by the download
    ctx.close()


def test_pipelined_upload_spmv_download(gpu, port):
    """upload(k+1) / SpMV(k) / download(k-1) overlap on three streams: every result still exact."""
    rows, cols, indptr, indices, data = matgen.rmat_csr(30000, 900000, 23)
    words = port.quantize(data)
    ctx = capi.Context(0, capi.IMPL_FIXED)
    ctx.upload_matrix_csr(rows, cols, indptr, indices, words)
    rng = np.random.default_rng(8)
    n = 12
    xs = [capi.PinnedArray(cols) for _ in range(n)]
    ys = [capi.PinnedArray(rows) for _ in range(n)]
    for k in range(n):
        xs[k].array[:] = port.quantize(rng.random(cols, dtype=np.float32))
    for k in range(n):
        ctx.upload_vector(xs[k].array)
        ctx.spmv()
        ctx.download_result_async(ys[k].array)
    ctx.sync()
    for k in range(n):
        assert np.array_equal(ys[k].array, port.spmv_q824(indptr, indices, words, xs[k].array)), k
    # several SpMVs on one vector, then a new vector without an intervening download
    ctx.spmv(); ctx.spmv()
    ctx.upload_vector(xs[3].array)
    ctx.spmv()
    assert np.array_equal(ctx.download_result(), port.spmv_q824(indptr, indices, words, xs[3].array))
    ctx.close()


# ------------------------------------------------------------------------------------------
# fixtures produced by the reference C simulation
# ------------------------------------------------------------------------------------------

def test_deferred_download_state_machine(gpu, port):
    """Downloads requested right after an SpMV are deferred onto the next launch (flag pipeline, two device
    y buffers). Every way out of the deferred state must deliver the right vector: the next whole-matrix
    SpMV, a row-partition launch, hsb_sync, the blocking download, and a second request in a row."""
    rows, cols, indptr, indices, data = matgen.rmat_csr(3000, 60000, 21)
    r2, c2, ip2 = matgen.pad_csr(rows, cols, indptr, 128, 8)
    words = port.quantize(data)
    rng = np.random.default_rng(4)
    xs = [port.quantize(rng.random(c2, dtype=np.float32)) for _ in range(6)]
    wants = [port.spmv_q824(ip2, indices, words, x) for x in xs]
    rpp = 1024
    ctx = capi.Context(0, capi.IMPL_FIXED)
    ctx.upload_matrix_csr(r2, c2, ip2, indices, words, rpp)
    nparts = (r2 + rpp - 1) // rpp
    out = [capi.PinnedArray(r2) for _ in range(8)]

    def all_partitions(x):
        ctx.upload_vector(x)
        for j in range(nparts):
            rows_here = min(rpp, r2 - j * rpp)
            ctx.spmv_row_partition(j, rows_here // 16, 1, nparts, c2)

    # deferred -> consumed by the next whole-matrix SpMV (twice in a row) -> resolved by hsb_sync
    for k in range(3):
        ctx.upload_vector(xs[k]); ctx.spmv(); ctx.download_result_async(out[k].array)
    ctx.sync()
    for k in range(3):
        assert np.array_equal(out[k].array, wants[k]), k
    # the same into pageable memory (copy engine + second device y buffer instead of the drain-to-host path),
    # and with the drain-to-host path switched off
    pageable = [np.zeros(r2, np.uint32) for _ in range(4)]
    for k in range(4):
        ctx.upload_vector(xs[k]); ctx.spmv(); ctx.download_result_async(pageable[k])
    ctx.sync()
    for k in range(4):
        assert np.array_equal(pageable[k], wants[k]), k
    ctx.set_option("host_drain", 0)
    for k in range(4):
        ctx.upload_vector(xs[k + 1]); ctx.spmv(); ctx.download_result_async(out[k].array)
    ctx.sync()
    for k in range(4):
        assert np.array_equal(out[k].array, wants[k + 1]), k
    ctx.set_option("host_drain", 1)
    ctx.set_option("flags", 0)
    for k in range(4):
        ctx.upload_vector(xs[k]); ctx.spmv(); ctx.download_result_async(out[k].array)
    ctx.sync()
    for k in range(4):
        assert np.array_equal(out[k].array, wants[k]), k
    ctx.set_option("flags", 1)
    # deferred -> a row-partition launch comes next (immediate resolution), then the partition flow's own result
    ctx.upload_vector(xs[3]); ctx.spmv(); ctx.download_result_async(out[3].array)
    all_partitions(xs[4])
    ctx.download_result_async(out[4].array)
    ctx.sync()
    assert np.array_equal(out[3].array, wants[3]) and np.array_equal(out[4].array, wants[4])
    # deferred -> blocking download of the same result; two requests for one result; device y after sync
    ctx.upload_vector(xs[5]); ctx.spmv()
    ctx.download_result_async(out[5].array)
    ctx.download_result_async(out[6].array)
    y = ctx.download_result()
    ctx.sync()
    assert np.array_equal(y, wants[5]) and np.array_equal(out[5].array, wants[5]) and np.array_equal(out[6].array, wants[5])
    # a long pipelined run alternating two vectors and two host buffers (what hsb_time_e2e issues)
    px = [capi.PinnedArray(c2) for _ in range(2)]
    px[0].array[:] = xs[0]; px[1].array[:] = xs[1]
    sec = ctx.time_e2e([px[0].array, px[1].array], [out[0].array, out[1].array], 200, async_download=True)
    assert sec > 0
    assert np.array_equal(out[0].array, wants[0]) and np.array_equal(out[1].array, wants[1])
    sec = ctx.time_e2e([px[0].array, px[1].array], [out[0].array, out[1].array], 20, async_download=False)
    assert np.array_equal(out[0].array, wants[0]) and np.array_equal(out[1].array, wants[1])
    ctx.close()



def test_overlapping_small_launches(gpu, port):
    """Small matrices use a handful of CTAs, so several consecutive launches are resident at the same time
    (a launch neither waits for its predecessor before working nor holds back its successor): the
    accumulator rotation, its reuse guard, the x flags and the end-of-launch drains must still give exact
    results, launch after launch."""
    for rows_, nnz_, rpp in ((600, 9000, 0), (3000, 40000, 1024), (20000, 300000, 0)):
        rows, cols, indptr, indices, data = matgen.rmat_csr(rows_, nnz_, 31)
        r2, c2, ip2 = matgen.pad_csr(rows, cols, indptr, 128, 8)
        words = port.quantize(data)
        rng = np.random.default_rng(9)
        xs = [port.quantize(rng.random(c2, dtype=np.float32)) for _ in range(2)]
        wants = [port.spmv_q824(ip2, indices, words, x) for x in xs]
        ctx = capi.Context(0, capi.IMPL_FIXED)
        ctx.upload_matrix_csr(r2, c2, ip2, indices, words, rpp)
        px = [capi.PinnedArray(c2) for _ in range(2)]
        py = [capi.PinnedArray(r2) for _ in range(2)]
        for k in range(2):
            px[k].array[:] = xs[k]
        xh, yh = [b.array for b in px], [b.array for b in py]
        for rep in range(3):
            ctx.time_e2e(xh, yh, 300, async_download=True)
            assert np.array_equal(yh[0], wants[0]) and np.array_equal(yh[1], wants[1]), (rows_, rep)
        # many launches on one vector, then the other, results only at the end
        for k in (0, 1, 0, 1):
            ctx.upload_vector(xh[k])
            for _ in range(37):
                ctx.spmv()
            assert np.array_equal(ctx.download_result(), wants[k]), (rows_, k)
        if rpp:
            nparts = (r2 + rpp - 1) // rpp
            for it in range(20):
                k = it & 1
                ctx.upload_vector(xh[k])
                for j in range(nparts):
                    ctx.spmv_row_partition(j, min(rpp, r2 - j * rpp) // 16, 1, nparts, c2)
                ctx.download_result_async(yh[k])
            ctx.sync()
            assert np.array_equal(yh[0], wants[0]) and np.array_equal(yh[1], wants[1])
        # float arithmetic through the same machinery
        ctx.close()
    rows, cols, indptr, indices, data = matgen.rmat_csr(3000, 40000, 33)
    r2, c2, ip2 = matgen.pad_csr(rows, cols, indptr, 128, 8)
    x = np.zeros(c2, np.float32)
    x[:cols] = np.random.default_rng(2).random(cols, dtype=np.float32)
    ctx = capi.Context(0, capi.IMPL_FLOAT_POB)
    ctx.upload_matrix_csr(r2, c2, ip2, indices, data)
    ctx.upload_vector(x)
    for _ in range(50):
        ctx.spmv()
    check_float(ctx.download_result(), port, ip2, indices, data, x)
    ctx.close()


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
@pytest.mark.parametrize("impl", hsoracle.IMPLS)
def test_reference_fixture_through_cpsr_images(gpu, port, path, impl):
    """CSR -> reference channel images (oracle restatement of the reference host code) ->
    hsb_upload_matrix_cpsr -> hsb_spmv_row_partition -> the y the reference csim produced."""
    g = np.load(path)
    rows, cols = int(g["rows"]), int(g["cols"])
    indptr, indices, skip = g["indptr"], g["indices"], bool(g["skip"])
    cfg = capi.get_config(capi.IMPL_BY_NAME[impl])
    IF, OB, VB = cfg.interleave_factor, cfg.logical_ob_size, cfg.logical_vb_size
    data = g["data_" + impl] if ("data_" + impl) in g else g["data"]
    x = g["x_" + impl] if ("x_" + impl) in g else g["x"]
    r2, c2, ip2 = matgen.pad_csr(rows, cols, indptr, 128 * IF, 8)
    assert r2 == int(g[impl + "_rows_padded"])
    xpad = np.zeros(c2, np.float32)
    xpad[:cols] = x
    if impl == "fixed":
        words, xw, kind = port.quantize(data), port.quantize(xpad), hsoracle.VAL_Q824
    else:
        words, xw, kind = data.view(np.uint32), xpad.view(np.uint32), hsoracle.VAL_FLOAT_BITS
    m = port.csr2cpsr(r2, c2, ip2, indices, words, 8, OB, VB, 16 * IF, skip, kind)
    images = m.channel_images(IF)
    ctx = capi.Context(0, impl)
    ctx.upload_matrix_cpsr(images, m.n_row_parts, m.n_col_parts, r2, c2)
    ctx.upload_vector(xw)
    for rp in range(m.n_row_parts):
        rows_here = OB if (rp < m.n_row_parts - 1 or r2 % OB == 0) else r2 % OB
        ctx.spmv_row_partition(rp, rows_here // 16, m.n_col_parts, m.n_col_parts * m.n_row_parts, c2)
    y = ctx.download_result()
    ctx.close()
    if impl == "fixed":
        assert np.array_equal(y, g["fixed_y"])
    else:
        check_float(y, port, ip2, indices, data, xpad)
        # and as close to the reference simulation's own float result
        _, sa = port.spmv_f64(ip2, indices, data, xpad)
        yr = g[impl + "_y"].view(np.float32).astype(np.float64)
        assert np.all(np.abs(y.view(np.float32) - yr) <= 2 * TOL * sa + 1e-30)


# ------------------------------------------------------------------------------------------
# float paths
# ------------------------------------------------------------------------------------------
FLOAT_CASES = [
    ("rand_4096_1pct", lambda: matgen.random_csr(4096, 4096, 0.01, 0xC0FFEE01, values="u01")),
    ("transformer_like", lambda: matgen.bernoulli_csr(512, 33288, 0.05, 0xC0FFEE03)),
    ("rmat_20000", lambda: matgen.rmat_csr(20000, 600000, 8, values="normal")),
]


@pytest.mark.parametrize("impl", ["float_pob", "float_stall"])
@pytest.mark.parametrize("name,make", FLOAT_CASES, ids=[c[0] for c in FLOAT_CASES])
def test_float_within_tolerance(gpu, port, impl, name, make):
    rows, cols, indptr, indices, data = make()
    x = (np.random.default_rng(2).random(cols, dtype=np.float32) * 2 - 1).astype(np.float32)
    ctx = capi.Context(0, impl)
    ctx.upload_matrix_csr(rows, cols, indptr, indices, data)
    ctx.upload_vector(x)
    ctx.spmv()
    y = ctx.download_result()
    ctx.close()
    check_float(y, port, indptr, indices, data, x)


# ------------------------------------------------------------------------------------------
# drop-in top_wrapper against the reference's own top_wrapper
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("impl", hsoracle.IMPLS)
def test_top_wrapper_drop_in(gpu, port, impl):
    if not hsoracle.ref_available(impl):
        pytest.skip("oracle/_ref not present")
    ref = hsoracle.Ref(impl)
    IF, OB, VB = ref.INTERLEAVE_FACTOR, ref.LOGICAL_OB_SIZE, ref.LOGICAL_VB_SIZE
    rows, cols, indptr, indices, data = matgen.rmat_csr(3000, 40000, 31)
    if impl != "fixed":
        data = (data - np.float32(0.5)).astype(np.float32)
    r2, c2, ip2 = matgen.pad_csr(rows, cols, indptr, 128 * IF, 8)
    x = np.random.default_rng(6).random(c2, dtype=np.float32)
    words = ref.val_from_float(data)
    xw = ref.val_from_float(x)
    kind = hsoracle.VAL_Q824 if impl == "fixed" else hsoracle.VAL_FLOAT_BITS
    m = port.csr2cpsr(r2, c2, ip2, indices, words, 8, OB, VB, 16 * IF, True, kind)
    images = m.channel_images(IF)
    y_ref = np.zeros(r2, np.uint32)
    y_gpu = np.zeros(r2, np.uint32)
    ref.top_wrapper(images, xw, y_ref, 0, r2 // 16, m.n_col_parts, m.n_col_parts, c2)
    capi.top_wrapper(impl, images, xw, y_gpu, 0, r2 // 16, m.n_col_parts, m.n_col_parts, c2)
    if impl == "fixed":
        assert np.array_equal(y_gpu, y_ref)
    else:
        _, sa = port.spmv_f64(ip2, indices, data, x)
        d = np.abs(y_gpu.view(np.float32).astype(np.float64) - y_ref.view(np.float32).astype(np.float64))
        assert np.all(d <= 2 * TOL * sa + 1e-30)


# ------------------------------------------------------------------------------------------
# full BASELINE size (config C2 stand-in): bit-exact against the closed form + properties
# ------------------------------------------------------------------------------------------
def test_fixed_googleplus_size(gpu, port):
    rows, cols, indptr, indices, data = matgen.rmat_csr(107614, 13_670_000, 0xC0FFEE02)
    data = (data * np.float32(0.05)).astype(np.float32)
    data[: indptr[1]] = 200.0                      # make row 0 saturate if it is not empty
    r2, c2, ip2 = matgen.pad_csr(rows, cols, indptr, 128, 8)
    words = port.quantize(data)
    rng = np.random.default_rng(3)
    x1 = np.zeros(c2, np.float32)
    x1[:cols] = rng.integers(0, 2, cols).astype(np.float32)          # reference-style {0,1}
    ctx = capi.Context(0, capi.IMPL_FIXED)
    ctx.upload_matrix_csr(r2, c2, ip2, indices, words)
    xw = port.quantize(x1)
    ctx.upload_vector(xw)
    ctx.spmv()
    y1 = ctx.download_result()
    assert np.array_equal(y1, port.spmv_q824(ip2, indices, words, xw))
    # idempotence: the same launch again gives the same bits (accumulators are re-zeroed)
    ctx.spmv()
    assert np.array_equal(ctx.download_result(), y1)
    # monotonicity property of the unsigned fixed path: x <= x' element-wise  =>  y <= y'
    xw2 = port.quantize(np.minimum(x1 + rng.random(c2, dtype=np.float32), 255).astype(np.float32))
    ctx.upload_vector(xw2)
    ctx.spmv()
    y2 = ctx.download_result()
    assert np.all(y2 >= y1)
    assert np.array_equal(y2, port.spmv_q824(ip2, indices, words, xw2))
    ctx.close()


# ------------------------------------------------------------------------------------------
# the C++ host drivers (hisparse_b200/host/host.cpp == the reference's sw/host.cpp flow)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("impl", hsoracle.IMPLS)
def test_cpp_host_driver(gpu, impl):
    import subprocess
    host = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hisparse_b200", "host")
    subprocess.run(["make", "-s", "-C", host], check=True, stdout=subprocess.DEVNULL)
    r = subprocess.run([os.path.join(host, "bin", "host_" + impl)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "===== All Test Passed! =====" in r.stdout


# ------------------------------------------------------------------------------------------
# GPU-side preprocessing (hsb_upload_matrix_csr_gpu): same results, structurally valid format
# ------------------------------------------------------------------------------------------
GPU_FORMAT_CASES = [
    ("empty", lambda: (128, 64, np.zeros(129, np.uint32), np.zeros(0, np.uint32), np.zeros(0, np.float32)), 0),
    ("dense128", lambda: matgen.dense_csr(128, 128), 0),
    ("uniform_unsorted_cols", lambda: matgen.uniform_sparse_csr(1000, 1024, 10), 0),
    ("multi_tile_70000", lambda: matgen.random_csr(900, 70000, 0.003, 7), 0),
    ("rmat_row_partitions", lambda: matgen.rmat_csr(20000, 300000, 9), 4096),
    ("rmat_60000", lambda: matgen.rmat_csr(60000, 2500000, 12), 0),
    ("one_long_row", lambda: (4, 70000, np.array([0, 0, 70000, 70000, 70001], np.uint32),
                              np.concatenate([np.arange(70000), [3]]).astype(np.uint32),
                              np.full(70001, 0.001, np.float32)), 0),
]


@pytest.mark.parametrize("name,make,rpp", GPU_FORMAT_CASES, ids=[c[0] for c in GPU_FORMAT_CASES])
def test_gpu_built_format(gpu, port, name, make, rpp):
    rows, cols, indptr, indices, data = make()
    words = port.quantize(data)
    xw = port.quantize(np.random.default_rng(1).random(cols, dtype=np.float32))
    ctx = capi.Context(0, capi.IMPL_FIXED)
    ctx.upload_matrix_csr(rows, cols, indptr, indices, words, rpp, on_gpu=True)
    # 1. the device-built format decodes back to the CSR it came from
    f = capi.Format.from_context(ctx)
    ip, ix, vv = f.expand()
    assert np.array_equal(ip, indptr)
    r = np.repeat(np.arange(rows, dtype=np.int64), np.diff(indptr.astype(np.int64)))
    a = np.lexsort((words, indices, r))
    b = np.lexsort((vv, ix, r))
    assert np.array_equal(indices[a], ix[b]) and np.array_equal(words[a], vv[b])
    # 2. and the SpMV over it is bit-exact
    ctx.upload_vector(xw)
    if rpp:
        nparts = (rows + rpp - 1) // rpp
        for j in range(nparts):
            ctx.spmv()          # whole-matrix launches and ...
        ctx.spmv()
    ctx.spmv()
    assert np.array_equal(ctx.download_result(), port.spmv_q824(indptr, indices, words, xw))
    # 3. same geometry as the host builder (streams, slices, padded slots)
    sg, sh = ctx.stats(), capi.Format(rows, cols, indptr, indices, words, rpp).stats()
    for k in ("n_streams", "n_slices", "n_elems", "n_col_tiles", "tile_cols", "n_row_parts"):
        assert sg[k] == sh[k], k
    ctx.close()


def test_cpsr_upload_rejects_truncated_image(gpu, port):
    cfg = capi.get_config(0)
    rows, cols, indptr, indices, data = matgen.random_csr(256, 512, 0.05, 9)
    m = port.csr2cpsr(rows, cols, indptr, indices, port.quantize(data), 8, cfg.logical_ob_size,
                      cfg.logical_vb_size, 16, False, hsoracle.VAL_Q824)
    images = m.channel_images(1)
    images[3] = images[3][:-1]
    ctx = capi.Context(0, capi.IMPL_FIXED)
    with pytest.raises(capi.HsbError):
        ctx.upload_matrix_cpsr(images, 1, 1, rows, cols)
    ctx.close()


def test_cpsr_images_large_float_stall(gpu, port):
    """float_stall: 8-way interleaved virtual channels, rows rounded to 1024, skip-empty-rows markers,
    three column partitions -- through the on-device CPSR decoder."""
    cfg = capi.get_config(capi.IMPL_FLOAT_STALL)
    IF = cfg.interleave_factor
    rows, cols, indptr, indices, data = matgen.rmat_csr(70000, 1500000, 41, values="normal")
    r2, c2, ip2 = matgen.pad_csr(rows, cols, indptr, 128 * IF, 8)
    x = np.zeros(c2, np.float32)
    x[:cols] = np.random.default_rng(2).random(cols, dtype=np.float32) * 2 - 1
    m = port.csr2cpsr(r2, c2, ip2, indices, data.view(np.uint32), 8, cfg.logical_ob_size, cfg.logical_vb_size,
                      16 * IF, True, hsoracle.VAL_FLOAT_BITS)
    ctx = capi.Context(0, capi.IMPL_FLOAT_STALL)
    ctx.upload_matrix_cpsr(m.channel_images(IF), m.n_row_parts, m.n_col_parts, r2, c2)
    ctx.upload_vector(x)
    ctx.spmv_row_partition(0, r2 // 16, m.n_col_parts, m.n_col_parts * m.n_row_parts, c2)
    check_float(ctx.download_result(), port, ip2, indices, data, x)
    ctx.close()


# ------------------------------------------------------------------------------------------
# iterative callers: x <- alpha (*) A x (+) beta on the device
# ------------------------------------------------------------------------------------------
def _pagerank_matrix(n, nnz, seed):
    """column-stochastic-ish link matrix: value = 1 / out-degree of the source column"""
    rows, cols, indptr, indices, _ = matgen.rmat_csr(n, nnz, seed)
    outdeg = np.maximum(np.bincount(indices, minlength=cols), 1)
    data = (1.0 / outdeg[indices]).astype(np.float32)
    r2, c2, ip2 = matgen.pad_csr(rows, cols, indptr, 128, 128)
    return r2, c2, ip2, indices, data


def test_iterate_fixed_bit_exact(gpu, port):
    from hisparse_b200 import sharding
    r2, c2, ip2, indices, data = _pagerank_matrix(6000, 90000, 41)
    assert r2 == c2
    words = port.quantize(data)
    alpha, beta = int(port.quantize(np.float32([0.85]))[0]), int(port.quantize(np.float32([0.15 / 8]))[0])
    x0 = port.quantize(np.full(c2, 1.0 / 8, np.float32))
    x = x0.copy()
    for _ in range(7):
        x = hsoracle.axpb_q824(alpha, port.spmv_q824(ip2, indices, words, x), beta)
    want = port.spmv_q824(ip2, indices, words, x)
    ctx = capi.Context(0, capi.IMPL_FIXED)
    ctx.upload_matrix_csr(r2, c2, ip2, indices, words)
    ctx.upload_vector(x0)
    ctx.iterate(7, alpha, beta)
    ctx.spmv()
    assert np.array_equal(ctx.download_result(), want)
    # step by step, with downloads in between and an upload restarting the iteration
    ctx.upload_vector(x0)
    x = x0.copy()
    for k in range(3):
        ctx.spmv()
        y = ctx.download_result()
        assert np.array_equal(y, port.spmv_q824(ip2, indices, words, x)), k
        ctx.axpb_to_vector(alpha, beta, 0)
        ctx.vector_commit()
        x = hsoracle.axpb_q824(alpha, y, beta)
    ctx.spmv()
    assert np.array_equal(ctx.download_result(), port.spmv_q824(ip2, indices, words, x))
    ctx.close()


def test_iterate_float(gpu, port):
    r2, c2, ip2, indices, data = _pagerank_matrix(6000, 90000, 43)
    alpha, beta = np.float32(0.85), np.float32(0.15 / r2)
    x = np.full(c2, 1.0 / r2, np.float64)
    x0 = x.astype(np.float32)
    for _ in range(6):
        y64, _ = port.spmv_f64(ip2, indices, data, x.astype(np.float32))
        x = float(alpha) * y64 + float(beta)
    y64, sa = port.spmv_f64(ip2, indices, data, x.astype(np.float32))
    ctx = capi.Context(0, capi.IMPL_FLOAT_POB)
    ctx.upload_matrix_csr(r2, c2, ip2, indices, data)
    ctx.upload_vector(x0)
    ctx.iterate(6, int(alpha.view(np.uint32)), int(beta.view(np.uint32)))
    ctx.spmv()
    y = ctx.download_result().view(np.float32).astype(np.float64)
    ctx.close()
    # rounding differences compound over the iterations: 1e-4 of the row's absolute sum after 7 SpMVs
    assert np.all(np.abs(y - y64) <= 1e-4 * sa + 1e-12)


def test_iterate_two_row_block_shards(gpu, port):
    """The multi-GPU iteration on one device: two contexts hold the two row-block shards, each writes its
    slice of the next vector at its row offset (hsb_axpb_to_vector), the slices are exchanged (the
    all-gather of tests/pagerank.py) and committed. Bit-equal to the single-context iteration."""
    import torch
    from hisparse_b200 import sharding
    r2, c2, ip2, indices, data = _pagerank_matrix(5000, 70000, 45)
    words = port.quantize(data)
    alpha, beta = int(port.quantize(np.float32([0.85]))[0]), int(port.quantize(np.float32([0.02]))[0])
    x0 = port.quantize(np.full(c2, 0.125, np.float32))
    x = x0.copy()
    for _ in range(4):
        x = hsoracle.axpb_q824(alpha, port.spmv_q824(ip2, indices, words, x), beta)
    want = port.spmv_q824(ip2, indices, words, x)
    bounds = sharding.shard_bounds(ip2, 2)
    ctxs = []
    for g in range(2):
        sip, six, sw = sharding.extract_shard(ip2, indices, words, bounds[g], bounds[g + 1])
        c = capi.Context(0, capi.IMPL_FIXED)
        c.upload_matrix_csr(bounds[g + 1] - bounds[g], c2, sip, six, sw)
        c.upload_vector(x0)
        ctxs.append(c)

    class Dev:
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 2}

    for _ in range(4):
        for g, c in enumerate(ctxs):
            c.spmv()
            c.axpb_to_vector(alpha, beta, bounds[g])
            c.sync()
        nxt = [torch.as_tensor(Dev(c.device_x_next(), c2), device="cuda:0") for c in ctxs]
        nxt[0][bounds[1]:bounds[2]] = nxt[1][bounds[1]:bounds[2]]
        nxt[1][bounds[0]:bounds[1]] = nxt[0][bounds[0]:bounds[1]]
        torch.cuda.synchronize()
        for c in ctxs:
            c.vector_commit()
    got = []
    for c in ctxs:
        c.spmv()
        got.append(c.download_result())
        c.close()
    assert np.array_equal(np.concatenate(got), want)


def test_peer_iteration_single_rank(gpu, port):
    """hsb_peer_export / hsb_peer_connect / hsb_axpb_to_peers with a world of one (the multi-GPU runs are
    tests/pagerank.py --p2p --check under torchrun): same kernel, same arrival-flag wait, bit-exact."""
    from hisparse_b200 import sharding
    r2, c2, ip2, indices, data = _pagerank_matrix(6000, 90000, 47)
    words = port.quantize(data)
    alpha, beta = int(port.quantize(np.float32([0.85]))[0]), int(port.quantize(np.float32([0.02]))[0])
    x0 = port.quantize(np.full(c2, 0.125, np.float32))
    x = x0.copy()
    for _ in range(9):
        x = hsoracle.axpb_q824(alpha, port.spmv_q824(ip2, indices, words, x), beta)
    ctx = capi.Context(0, capi.IMPL_FIXED)
    ctx.upload_matrix_csr(r2, c2, ip2, indices, words)
    ctx.upload_vector(x0)
    ctx.peer_connect(1, 0, ctx.peer_export())
    for _ in range(9):
        ctx.spmv()
        ctx.axpb_to_peers(alpha, beta, 0)
        ctx.vector_commit()
    ctx.spmv()
    assert np.array_equal(ctx.download_result(), port.spmv_q824(ip2, indices, words, x))
    ctx.close()


def test_cpp_benchmark_driver(gpu):
    """hisparse_b200/host/benchmark.cpp (the mirror of sw/benchmark.cpp) on a synthetic matrix: prints the
    reference's result line and the host-buffer pipeline line."""
    import subprocess
    host = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hisparse_b200", "host")
    subprocess.run(["make", "-s", "-C", host], check=True, stdout=subprocess.DEVNULL)
    r = subprocess.run([os.path.join(host, "bin", "benchmark_fixed"), "rmat:20000:400000:3"], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "GOPS }" in r.stdout and "with host buffers" in r.stdout and "===== Benchmark Finished =====" in r.stdout


# ------------------------------------------------------------------------------------------
# row-block shards + gather of y fused into the result drain (hsb_gather_*), on one device:
# the contexts of one process reach each other's buffers through plain pointers, the kernels,
# the arrival flags and the bookkeeping are the ones the multi-GPU runs use (bench.py `sharded`)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("world,all_targets", [(2, False), (3, True)])
def test_sharded_spmv_gather_of_y(gpu, port, world, all_targets):
    from hisparse_b200 import sharding
    rows, cols, indptr, indices, data = matgen.rmat_csr(30000, 900000, 77)
    data = (data * np.float32(0.05)).astype(np.float32)
    r2, c2, ip2 = matgen.pad_csr(rows, cols, indptr, 128, 8)
    words = port.quantize(data)
    rng = np.random.default_rng(5)
    xs = [port.quantize(rng.random(c2, dtype=np.float32)) for _ in range(3)]
    bounds = sharding.shard_bounds(ip2, world)
    ctxs = []
    for g in range(world):
        sip, six, sw = sharding.extract_shard(ip2, indices, words, bounds[g], bounds[g + 1])
        c = capi.Context(0, capi.IMPL_FIXED)
        c.upload_matrix_csr(bounds[g + 1] - bounds[g], c2, sip, six, sw)
        c.upload_vector(xs[0])
        ctxs.append(c)
    blobs = np.concatenate([c.gather_export(r2, want_buffer=(all_targets or g == 0)) for g, c in enumerate(ctxs)])
    for g, c in enumerate(ctxs):
        c.gather_connect(world, g, bounds[g], blobs)
    targets = ctxs if all_targets else ctxs[:1]
    # one SpMV, drained (and gathered) by hsb_sync
    for c in ctxs:
        c.spmv()
    for c in ctxs:
        c.sync()
    want = port.spmv_q824(ip2, indices, words, xs[0])
    for t in targets:
        assert np.array_equal(t.download_gathered(), want)
    # back-to-back SpMVs: the first is drained by the END of the second launch, the second by hsb_sync;
    # the gathered vector holds the last one
    for k in (1, 2):
        for c in ctxs:
            c.upload_vector(xs[k])
            c.spmv()
    for c in ctxs:
        c.sync()
    want = port.spmv_q824(ip2, indices, words, xs[2])
    for t in targets:
        assert np.array_equal(t.download_gathered(), want)
    # every rank still has its own block
    for g, c in enumerate(ctxs):
        assert np.array_equal(c.download_result(), want[bounds[g]:bounds[g + 1]])
    if not all_targets:
        with pytest.raises(capi.HsbError):
            ctxs[1].download_gathered()
    for c in ctxs:
        c.close()


# ------------------------------------------------------------------------------------------
# full BASELINE sizes of the float configurations
# ------------------------------------------------------------------------------------------
def test_float_ogbl_ppa_size(gpu, port):
    """config C4 stand-in at full size (576,289^2, ~42 M non-zeros, 14 column tiles), fp32"""
    rows, cols, indptr, indices, data = matgen.rmat_csr(576289, 42_460_000, 0xC0FFEE04, symmetric=True, oversample=1.5)
    r2, c2, ip2 = matgen.pad_csr(rows, cols, indptr, 128, 8)
    x = np.zeros(c2, np.float32)
    x[:cols] = np.random.default_rng(4).random(cols, dtype=np.float32)
    ctx = capi.Context(0, "float_pob")
    ctx.upload_matrix_csr(r2, c2, ip2, indices, data)
    ctx.upload_vector(x)
    ctx.spmv()
    y = ctx.download_result()
    assert ctx.stats()["nnz"] == int(ip2[-1]) and ctx.stats()["n_col_tiles"] > 8
    check_float(y, port, ip2, indices, data, x)
    # linearity in x (up to rounding): A(2x) = 2 A x exactly in binary floating point
    ctx.upload_vector((2 * x).astype(np.float32))
    ctx.spmv()
    y2 = ctx.download_result().view(np.float32).astype(np.float64)
    _, sa = port.spmv_f64(ip2, indices, data, x)
    assert np.all(np.abs(y2 - 2 * y.view(np.float32).astype(np.float64)) <= 4 * TOL * sa + 1e-30)
    ctx.close()


@pytest.mark.parametrize("impl", ["float_pob", "float_stall"])
def test_float_transformer_50_percent(gpu, port, impl):
    """config C3 at the dense end of the series: 512 x 33,288, 50 % Bernoulli mask, N(0, 0.05^2) values, x in (-1, 1)"""
    rows, cols, indptr, indices, data = matgen.bernoulli_csr(512, 33288, 0.5, 0xC0FFEE03)
    IF = 8 if impl == "float_stall" else 1
    r2, c2, ip2 = matgen.pad_csr(rows, cols, indptr, 128 * IF, 8)
    x = (np.random.default_rng(2).random(c2, dtype=np.float32) * 2 - 1).astype(np.float32)
    ctx = capi.Context(0, impl)
    ctx.upload_matrix_csr(r2, c2, ip2, indices, data)
    ctx.upload_vector(x)
    ctx.spmv()
    y = ctx.download_result()
    ctx.close()
    assert int(ip2[-1]) > 8_000_000
    check_float(y, port, ip2, indices, data, x)


# ------------------------------------------------------------------------------------------
# the reference's OWN csim harness (spmv_csim/csim.cpp, unmodified, compiled where it lies) linked against
# the library: its top_wrapper symbol is replaced by hisparse_b200/host/top_wrapper.h (oracle/Makefile,
# csim_gpu_*). The harness formats with the reference's csr2cpsr, builds the 16 channel images, calls
# top_wrapper once per row partition and verifies against its own compute_ref (csim.cpp:203-381).
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("impl", hsoracle.IMPLS)
def test_reference_csim_harness_on_gpu(gpu, impl):
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "csim_gpu_" + impl)
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/csim_gpu_%s not built (needs /root/reference at build time)" % impl)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    out = r.stdout
    # main() (csim.cpp:597-614) runs test_basic, test_basic_sparse, test_large_sparse and then the dataset
    # tests, whose .npz files the reference does not ship: the cnpy stand-in aborts at the first of them
    assert out.count("INFO : Testcase passed.") == 3, out[-3000:] + r.stderr[-2000:]
    assert "Testcase failed" not in out and "top_wrapper on the GPU failed" not in r.stderr
    for name in ("on basic dense matrix", "on basic sparse matrix", "on uniform 100K 10"):
        assert "------ Running test: " + name in out
    assert "datasets are not available" in r.stderr


# ------------------------------------------------------------------------------------------
# narrow layout (hypersparse matrices; csrc/tile_format.h): chosen automatically for singleton-dominated
# matrices (tests/test_gpu_synth.py), FORCED here on ordinary ones, so that long rows (streams cut at 32),
# row partitions, partition launches and the pipelined host-buffer path all run through the narrow kernel
# ------------------------------------------------------------------------------------------
NARROW_GPU_CASES = [
    ("rmat_20000", lambda: matgen.rmat_csr(20000, 600000, 8), 0),
    ("multi_tile_70000", lambda: matgen.random_csr(900, 70000, 0.003, 7), 0),
    ("row_partitions", lambda: matgen.rmat_csr(20000, 300000, 9), 4096),
    ("dense128", lambda: matgen.dense_csr(128, 128), 0),
    ("one_long_row", lambda: (4, 70000, np.array([0, 0, 70000, 70000, 70001], np.uint32),
                              np.concatenate([np.arange(70000), [3]]).astype(np.uint32),
                              np.full(70001, 0.001, np.float32)), 0),
    ("singletons_92_tiles", lambda: matgen.random_csr(30000, 3000000, 0.000004, 13), 0),
]


@pytest.mark.parametrize("name,make,rpp", NARROW_GPU_CASES, ids=[c[0] for c in NARROW_GPU_CASES])
def test_narrow_layout_fixed_bit_exact(gpu, port, monkeypatch, name, make, rpp):
    monkeypatch.setenv("HSB_NARROW", "1")
    rows, cols, indptr, indices, data = make()
    words = port.quantize((data * np.float32(0.5)).astype(np.float32))
    rng = np.random.default_rng(3)
    xs = [port.quantize(rng.random(cols, dtype=np.float32)) for _ in range(3)]
    ctx = capi.Context(0, capi.IMPL_FIXED)
    ctx.upload_matrix_csr(rows, cols, indptr, indices, words, rpp)
    assert ctx.stats()["layout"] == 1
    # the device-built narrow format decodes back to the CSR, with the host builder's geometry
    ip, ix, vv = capi.Format.from_context(ctx).expand()
    assert np.array_equal(ip, indptr)
    r = np.repeat(np.arange(rows, dtype=np.int64), np.diff(indptr.astype(np.int64)))
    a, b = np.lexsort((words, indices, r)), np.lexsort((vv, ix, r))
    assert np.array_equal(indices[a], ix[b]) and np.array_equal(words[a], vv[b])
    sg, sh = ctx.stats(), capi.Format(rows, cols, indptr, indices, words, rpp).stats()
    for k in ("n_streams", "n_slices", "n_elems", "n_col_tiles", "layout"):
        assert sg[k] == sh[k], k
    # blocking, back-to-back (the end-of-launch drain), and one row partition at a time
    ctx.upload_vector(xs[0]); ctx.spmv()
    assert np.array_equal(ctx.download_result(), port.spmv_q824(indptr, indices, words, xs[0]))
    ctx.upload_vector(xs[1]); ctx.spmv(); ctx.spmv(); ctx.spmv()
    assert np.array_equal(ctx.download_result(), port.spmv_q824(indptr, indices, words, xs[1]))
    if rpp:
        ctx.upload_vector(xs[2])
        nparts = (rows + rpp - 1) // rpp
        for j in range(nparts):
            ctx.spmv_row_partition(j, min(rpp, rows - j * rpp) // 16, 1, nparts, cols + (-cols) % 8)
        assert np.array_equal(ctx.download_result(), port.spmv_q824(indptr, indices, words, xs[2]))
    ctx.close()


def test_narrow_layout_float(gpu, port, monkeypatch):
    monkeypatch.setenv("HSB_NARROW", "1")
    rows, cols, indptr, indices, data = matgen.rmat_csr(20000, 600000, 8, values="normal")
    x = (np.random.default_rng(2).random(cols, dtype=np.float32) * 2 - 1).astype(np.float32)
    ctx = capi.Context(0, "float_pob")
    ctx.upload_matrix_csr(rows, cols, indptr, indices, data)
    assert ctx.stats()["layout"] == 1
    ctx.upload_vector(x)
    ctx.spmv()
    y = ctx.download_result()
    ctx.close()
    check_float(y, port, indptr, indices, data, x)


@pytest.mark.parametrize("impl", ["fixed", "float_pob"])
def test_cpp_multi_gpu_driver_single_rank(gpu, impl):
    """hisparse_b200/host/benchmark_mgpu.cpp with a world of one (the N-GPU runs are `benchmark_mgpu_* <spec> N` on an
    N-GPU box): NCCL communicator, broadcast of x into the engine, gather connect, gathered y checked on the host."""
    import subprocess
    host = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hisparse_b200", "host")
    subprocess.run(["make", "-s", "-C", host], check=True, stdout=subprocess.DEVNULL)
    r = subprocess.run([os.path.join(host, "bin", "benchmark_mgpu_" + impl), "rmat:20000:400000:3", "1"], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "matches the host reference of the whole matrix" in r.stdout and "GOPS }" in r.stdout
    assert "===== Benchmark Finished =====" in r.stdout


def test_cpsr_ingestion_googleplus_size(gpu, port):
    """The reference's channel images of the C2-sized matrix (131 MB of 64-byte packets, 4 column partitions) handed
    over unchanged: decoded by one warp per lane-stream piece and formatted on the device. Bit-exact SpMV, and the
    whole upload (copy + decode + format) stays within a small multiple of the CSR route's time."""
    import time
    rows, cols, indptr, indices, data = matgen.rmat_csr(107614, 13_670_000, 0xC0FFEE02)
    r2, c2, ip2 = matgen.pad_csr(rows, cols, indptr, 128, 8)
    words = port.quantize((data * np.float32(0.05)).astype(np.float32))
    cfg = capi.get_config(capi.IMPL_FIXED)
    m = port.csr2cpsr(r2, c2, ip2, indices, words, 8, cfg.logical_ob_size, cfg.logical_vb_size, 16, True, hsoracle.VAL_Q824)
    images = m.channel_images(1)
    xw = port.quantize(np.random.default_rng(8).random(c2, dtype=np.float32))
    ctx = capi.Context(0, capi.IMPL_FIXED)
    ctx.upload_matrix_csr(r2, c2, ip2, indices, words)            # warm-up of the formatter's allocations
    t_csr, t_cpsr = [], []
    for _ in range(3):                                            # (uploads on a shared host jitter: the best of three counts)
        t0 = time.perf_counter()
        ctx.upload_matrix_csr(r2, c2, ip2, indices, words)
        t_csr.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        ctx.upload_matrix_cpsr(images, m.n_row_parts, m.n_col_parts, r2, c2)
        t_cpsr.append(time.perf_counter() - t0)
    t_csr, t_cpsr = min(t_csr), min(t_cpsr)
    assert ctx.stats()["nnz"] == indices.size
    ctx.upload_vector(xw)
    ctx.spmv()
    assert np.array_equal(ctx.download_result(), port.spmv_q824(ip2, indices, words, xw))
    ctx.close()
    print("CPSR images -> resident: %.1f ms; CSR -> resident: %.1f ms" % (1e3 * t_cpsr, 1e3 * t_csr))
    assert t_cpsr < 4 * t_csr + 0.05


def test_timed_out_flag_wait_skips_the_row_updates_and_is_reported(gpu, port):
    """A launch whose x never arrives (here: the slice of a second rank that never sends it) gives up after a bounded
    wait, makes NO row update, and every way of asking for the result reports the failure -- no hang, no stale y."""
    r2, c2, ip2, indices, data = _pagerank_matrix(3000, 40000, 51)
    words = port.quantize(data)
    alpha, beta = int(port.quantize(np.float32([0.85]))[0]), int(port.quantize(np.float32([0.02]))[0])
    x0 = port.quantize(np.full(c2, 0.125, np.float32))
    a, b = capi.Context(0, capi.IMPL_FIXED), capi.Context(0, capi.IMPL_FIXED)
    for c in (a, b):
        c.upload_matrix_csr(r2, c2, ip2, indices, words)
        c.upload_vector(x0)
    blobs = np.concatenate([a.peer_export(), b.peer_export()])
    a.peer_connect(2, 0, blobs)
    b.peer_connect(2, 1, blobs)
    a.spmv()
    a.axpb_to_peers(alpha, beta, 0)          # rank 0 sends its slice; rank 1 never does
    a.vector_commit()
    a.spmv()                                  # polls two arrival flags, one of which never comes
    with pytest.raises(capi.HsbError, match="gave up waiting"):
        a.sync()
    a.close()
    b.close()


def test_cpp_benchmark_driver_on_an_npz_dataset(gpu, tmp_path):
    """The reference's dataset route end to end: a scipy-style `.npz` CSR (compressed, int64
    index arrays, as datasets/download.sh delivers them) -> the C++ loader (own zlib reader in place of cnpy,
    sw/data_loader.h:51-70) -> benchmark.cpp's value reset, rounding and VAL_T conversion -> GPU; the driver prints the
    reference's result line and takes the CPSR route as well (same non-zero count through both)."""
    import subprocess
    rows, cols, indptr, indices, data = matgen.rmat_csr(30000, 700000, 61)
    path = str(tmp_path / "graph_30K_700K_csr_float32.npz")
    np.savez_compressed(path, shape=np.array([rows, cols], np.int64), data=data, indices=indices.astype(np.int64),
                        indptr=indptr.astype(np.int64), format=np.array("csr"))
    host = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hisparse_b200", "host")
    subprocess.run(["make", "-s", "-C", host], check=True, stdout=subprocess.DEVNULL)
    r = subprocess.run([os.path.join(host, "bin", "benchmark_fixed"), path], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "nnz %d," % indices.size in r.stdout and "same matrix" in r.stdout
    assert "GOPS }" in r.stdout and "===== Benchmark Finished =====" in r.stdout


def test_cpsr_upload_rejects_bad_columns_and_recovers(gpu, port):
    """the on-device decoder refuses a column id beyond the vector buffer / beyond the matrix (HSB_EINVAL, no partial
    state), and the context takes a well-formed upload afterwards"""
    cfg = capi.get_config(0)
    rows, cols, indptr, indices, data = matgen.random_csr(256, 512, 0.05, 9)
    words = port.quantize(data)
    m = port.csr2cpsr(rows, cols, indptr, indices, words, 8, cfg.logical_ob_size, cfg.logical_vb_size, 16, False,
                      hsoracle.VAL_Q824)
    good = m.channel_images(1)
    ctx = capi.Context(0, capi.IMPL_FIXED)
    for bad_col in (cfg.logical_vb_size + 5, 600):                 # beyond the vector buffer; inside it but beyond num_cols
        images = [im.copy() for im in good]
        im = images[2].reshape(-1, 16)
        n_hdr = 2                                                  # (1 + IF) header packets per partition, one partition
        r, k = np.argwhere(im[n_hdr:, :8] != 0xFFFFFFFF)[0]        # a real entry, not a marker
        im[n_hdr + r, k] = bad_col
        with pytest.raises(capi.HsbError, match="column"):
            ctx.upload_matrix_cpsr(images, 1, 1, rows, cols)
    ctx.upload_matrix_cpsr(good, 1, 1, rows, cols)
    xw = port.quantize(np.random.default_rng(4).random(cols, dtype=np.float32))
    ctx.upload_vector(xw)
    ctx.spmv()
    assert np.array_equal(ctx.download_result(), port.spmv_q824(indptr, indices, words, xw))
    ctx.close()


def test_two_contexts_interleaved_on_one_device(gpu, port):
    """two engines (fixed and float) on one GPU, calls interleaved: each keeps its own streams, flags and buffers"""
    rows, cols, indptr, indices, data = matgen.rmat_csr(8000, 200000, 71)
    words = port.quantize((data * np.float32(0.1)).astype(np.float32))
    rng = np.random.default_rng(6)
    a, b = capi.Context(0, "fixed"), capi.Context(0, "float_pob")
    a.upload_matrix_csr(rows, cols, indptr, indices, words)
    b.upload_matrix_csr(rows, cols, indptr, indices, data)
    for _ in range(3):
        xf = rng.random(cols, dtype=np.float32)
        xw = port.quantize(xf)
        a.upload_vector(xw); b.upload_vector(xf)
        a.spmv(); b.spmv(); a.spmv(); b.spmv()
        ya, yb = a.download_result(), b.download_result()
        assert np.array_equal(ya, port.spmv_q824(indptr, indices, words, xw))
        check_float(yb, port, indptr, indices, data, xf)
    a.close(); b.close()


def test_float_stall_images_through_the_narrow_layout(gpu, port, monkeypatch):
    """float_stall channel images (8-way interleave) decoded on the device and formatted into the NARROW layout"""
    monkeypatch.setenv("HSB_NARROW", "1")
    cfg = capi.get_config(capi.IMPL_FLOAT_STALL)
    IF = cfg.interleave_factor
    rows, cols, indptr, indices, data = matgen.rmat_csr(9000, 150000, 43, values="normal")
    r2, c2, ip2 = matgen.pad_csr(rows, cols, indptr, 128 * IF, 8)
    x = np.zeros(c2, np.float32)
    x[:cols] = np.random.default_rng(2).random(cols, dtype=np.float32) * 2 - 1
    m = port.csr2cpsr(r2, c2, ip2, indices, data.view(np.uint32), 8, cfg.logical_ob_size, cfg.logical_vb_size,
                      16 * IF, True, hsoracle.VAL_FLOAT_BITS)
    ctx = capi.Context(0, capi.IMPL_FLOAT_STALL)
    ctx.upload_matrix_cpsr(m.channel_images(IF), m.n_row_parts, m.n_col_parts, r2, c2)
    assert ctx.stats()["layout"] == 1
    ctx.upload_vector(x)
    ctx.spmv()
    check_float(ctx.download_result(), port, ip2, indices, data, x)
    ctx.close()


@pytest.mark.parametrize("base", [(1 << 32) - 40, (1 << 31) - 40, (3 << 32) - 7], ids=["2^32", "2^31", "3*2^32"])
def test_sequence_numbers_cross_a_32_bit_wrap(gpu, port, monkeypatch, base):
    """A context whose launch / upload / download numbers start just below a 32-bit wrap-around (HSB_DEBUG_SEQ_BASE):
    the flag words on the device are 32-bit and compared cyclically, the host keeps 64-bit numbers. The pipelined
    sequence, repeated launches on one vector and the iteration must all run across the wrap with exact results."""
    monkeypatch.setenv("HSB_DEBUG_SEQ_BASE", str(base))
    rows, cols, indptr, indices, data = matgen.rmat_csr(3000, 40000, 37)
    r2, c2, ip2 = matgen.pad_csr(rows, cols, indptr, 128, 8)
    words = port.quantize(data)
    rng = np.random.default_rng(12)
    ctx = capi.Context(0, capi.IMPL_FIXED)
    monkeypatch.delenv("HSB_DEBUG_SEQ_BASE")
    ctx.upload_matrix_csr(r2, c2, ip2, indices, words)
    n = 24
    xs = [capi.PinnedArray(c2) for _ in range(n)]
    ys = [capi.PinnedArray(r2) for _ in range(n)]
    for k in range(n):
        xs[k].array[:] = port.quantize(rng.random(c2, dtype=np.float32))
    wants = [port.spmv_q824(ip2, indices, words, b.array) for b in xs]
    for k in range(n):                                   # 24 uploads, launches and downloads: up to base + 24
        ctx.upload_vector(xs[k].array)
        ctx.spmv()
        ctx.download_result_async(ys[k].array)
    ctx.sync()
    for k in range(n):
        assert np.array_equal(ys[k].array, wants[k]), k
    for k in (5, 6):                                     # 2 x 37 launches on one vector: crosses the wrap
        ctx.upload_vector(xs[k].array)
        for _ in range(37):
            ctx.spmv()
        assert np.array_equal(ctx.download_result(), wants[k]), k
    ctx.time_e2e([xs[0].array, xs[1].array], [ys[0].array, ys[1].array], 200, async_download=True)
    assert np.array_equal(ys[0].array, wants[0]) and np.array_equal(ys[1].array, wants[1])
    ctx.time_e2e([xs[2].array, xs[3].array], [ys[2].array, ys[3].array], 50, async_download=False)
    assert np.array_equal(ys[2].array, wants[2]) and np.array_equal(ys[3].array, wants[3])
    ctx.close()


def test_contexts_release_their_device_memory(gpu, port):
    """create / upload / run / destroy twenty times (CSR, CPSR images, re-upload into a live context): free device
    memory ends where it started (within the allocator's granularity), i.e. no per-context or per-matrix leak"""
    import torch
    rows, cols, indptr, indices, data = matgen.rmat_csr(20000, 600000, 41)
    r2, c2, ip2 = matgen.pad_csr(rows, cols, indptr, 128, 8)
    words = port.quantize(data)
    x = port.quantize(np.random.default_rng(4).random(c2, dtype=np.float32))
    want = port.spmv_q824(ip2, indices, words, x)
    cfg = capi.get_config(capi.IMPL_FIXED)
    m = port.csr2cpsr(r2, c2, ip2, indices, words, 8, cfg.logical_ob_size, cfg.logical_vb_size, 16, True, hsoracle.VAL_Q824)

    def once(k):
        ctx = capi.Context(0, capi.IMPL_FIXED)
        if k % 3 == 2:
            ctx.upload_matrix_cpsr(m.channel_images(1), m.n_row_parts, m.n_col_parts, r2, c2)
        else:
            ctx.upload_matrix_csr(r2, c2, ip2, indices, words)
        if k % 3 == 1:
            ctx.upload_matrix_csr(r2, c2, ip2, indices, words)       # replaces the matrix of a live context
        ctx.set_replicas(2)
        ctx.upload_vector(x)
        ctx.spmv(); ctx.spmv()
        assert np.array_equal(ctx.download_result(), want), k
        ctx.close()

    once(0); once(1); once(2)                       # warm the allocator and the kernel images
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info(0)[0]
    for k in range(20):
        once(k)
    torch.cuda.synchronize()
    free1 = torch.cuda.mem_get_info(0)[0]
    assert free0 - free1 < (8 << 20), (free0, free1)


def test_two_host_threads_two_contexts(gpu, port):
    """two host threads, each driving its own context through the pipelined upload / SpMV / download sequence at the
    same time (ctypes drops the GIL inside the calls): per-context state only, the last-error string is per thread"""
    import threading
    mats = []
    for seed, rows_, nnz_ in ((51, 6000, 90000), (52, 15000, 400000)):
        rows, cols, indptr, indices, data = matgen.rmat_csr(rows_, nnz_, seed)
        r2, c2, ip2 = matgen.pad_csr(rows, cols, indptr, 128, 8)
        words = port.quantize(data)
        rng = np.random.default_rng(seed)
        xs = [port.quantize(rng.random(c2, dtype=np.float32)) for _ in range(2)]
        mats.append((r2, c2, ip2, indices, words, xs, [port.spmv_q824(ip2, indices, words, x) for x in xs]))
    errors = []

    def drive(t):
        try:
            r2, c2, ip2, indices, words, xs, wants = mats[t]
            ctx = capi.Context(0, capi.IMPL_FIXED)
            ctx.upload_matrix_csr(r2, c2, ip2, indices, words)
            px = [capi.PinnedArray(c2) for _ in range(2)]
            py = [capi.PinnedArray(r2) for _ in range(2)]
            for k in range(2):
                px[k].array[:] = xs[k]
            for rep in range(6):
                ctx.time_e2e([b.array for b in px], [b.array for b in py], 400, async_download=rep % 2 == 0)
                for k in range(2):
                    if not np.array_equal(py[k].array, wants[k]):
                        errors.append((t, rep, k))
                for k in range(2):
                    ctx.upload_vector(px[k].array); ctx.spmv()
                    if not np.array_equal(ctx.download_result(), wants[k]):
                        errors.append((t, rep, k, "sync"))
            # an error in this thread must not show up in the other one's hsb_last_error
            with pytest.raises(capi.HsbError):
                ctx.upload_vector(np.zeros(3, np.uint32))
            ctx.close()
        except Exception as e:                      # noqa: BLE001
            errors.append((t, repr(e)))

    th = [threading.Thread(target=drive, args=(t,)) for t in range(2)]
    for t in th: t.start()
    for t in th: t.join()
    assert not errors, errors


@pytest.mark.parametrize("narrow", [0, 1], ids=["wide", "narrow"])
@pytest.mark.parametrize("n,nnz", [(600, 9000), (6000, 90000), (60000, 1500000), (200000, 2400000)])   # the last one: several rows per thread in the update (128-bit path)
def test_iterate_one_launch_against_oracle_and_step_form(gpu, port, monkeypatch, narrow, n, nnz):
    """hsb_iterate as ONE cooperative launch (grid barriers between the SpMV and the update of x): the vector after k
    iterations, y of the last iteration and the following SpMV are bit-equal to the oracle's iteration and to the
    launch-per-step form; odd and even iteration counts (the two x buffers swap), calls mixed with uploads, downloads
    and plain SpMVs, both layouts."""
    monkeypatch.setenv("HSB_NARROW", str(narrow))
    r2, c2, ip2, indices, data = _pagerank_matrix(n, nnz, 61 + narrow)
    words = port.quantize(data)
    alpha, beta = int(port.quantize(np.float32([0.85]))[0]), int(port.quantize(np.float32([0.15 / 8]))[0])
    x0 = port.quantize(np.full(c2, 1.0 / 8, np.float32))
    ctx = capi.Context(0, capi.IMPL_FIXED)
    ctx.upload_matrix_csr(r2, c2, ip2, indices, words)
    assert ctx.stats()["layout"] == narrow
    ref = capi.Context(0, capi.IMPL_FIXED)
    ref.upload_matrix_csr(r2, c2, ip2, indices, words)
    ref.set_option("iterate_persistent", 0)
    x = x0.copy()
    done = 0
    ctx.upload_vector(x0); ref.upload_vector(x0)
    for k in (1, 2, 5, 4):
        ctx.iterate(k, alpha, beta); ref.iterate(k, alpha, beta)
        for _ in range(k):
            y = port.spmv_q824(ip2, indices, words, x)
            x = hsoracle.axpb_q824(alpha, y, beta)
        done += k
        got_y = ctx.download_result()
        assert np.array_equal(got_y, y), (done, "y of the last iteration")
        assert np.array_equal(ref.download_result(), y), (done, "step form")
        ctx.spmv(); ref.spmv()
        want = port.spmv_q824(ip2, indices, words, x)
        assert np.array_equal(ctx.download_result(), want), (done, "SpMV on the iterated vector")
        assert np.array_equal(ref.download_result(), want), (done, "step form")
    # a fresh upload restarts the iteration; a pinned, deferred download sits between the calls
    py = capi.PinnedArray(r2)
    ctx.upload_vector(x0)
    ctx.spmv()
    ctx.download_result_async(py.array)
    ctx.iterate(3, alpha, beta)
    ctx.sync()
    assert np.array_equal(py.array, port.spmv_q824(ip2, indices, words, x0))
    x = x0.copy()
    for _ in range(3):
        y = port.spmv_q824(ip2, indices, words, x)
        x = hsoracle.axpb_q824(alpha, y, beta)
    assert np.array_equal(ctx.download_result(), y)
    ctx.spmv()                                            # the step form continues from the state the launch left
    ctx.axpb_to_vector(alpha, beta, 0)
    ctx.vector_commit()
    ctx.spmv()
    x = hsoracle.axpb_q824(alpha, port.spmv_q824(ip2, indices, words, x), beta)
    assert np.array_equal(ctx.download_result(), port.spmv_q824(ip2, indices, words, x))
    ctx.close(); ref.close()


def test_iterate_one_launch_float_equals_step_form(gpu, port):
    """fp32: the cooperative launch and the launch-per-step form run the same arithmetic; row updates are atomic adds in
    no fixed order, so the two agree within the tolerance of the path (compounded over the iterations), not bit for bit"""
    r2, c2, ip2, indices, data = _pagerank_matrix(20000, 500000, 67)
    alpha, beta = np.float32(0.85), np.float32(0.15 / r2)
    x0 = np.full(c2, 1.0 / r2, np.float32)
    outs = []
    for persistent in (1, 0):
        ctx = capi.Context(0, capi.IMPL_FLOAT_POB)
        ctx.upload_matrix_csr(r2, c2, ip2, indices, data)
        ctx.set_option("iterate_persistent", persistent)
        ctx.upload_vector(x0)
        ctx.iterate(9, int(alpha.view(np.uint32)), int(beta.view(np.uint32)))
        outs.append(ctx.download_result().view(np.float32).astype(np.float64))
        ctx.close()
    x = x0.astype(np.float64)
    for _ in range(9):
        y64, sa = port.spmv_f64(ip2, indices, data, x.astype(np.float32))
        x = float(alpha) * y64 + float(beta)
    for y in outs:
        assert np.all(np.abs(y - y64) <= 1e-4 * sa + 1e-12)


def test_two_contexts_iterate_concurrently(gpu, port):
    """two host threads, two contexts, both inside hsb_iterate at the same time: each cooperative launch needs every SM
    for its grid barriers, so the two grids must never be resident half and half (the launches are gang-scheduled);
    both finish with exact results, no barrier time-out"""
    import threading
    r2, c2, ip2, indices, data = _pagerank_matrix(40000, 1200000, 71)
    words = port.quantize(data)
    alpha, beta = int(port.quantize(np.float32([0.85]))[0]), int(port.quantize(np.float32([0.15 / 8]))[0])
    x0 = port.quantize(np.full(c2, 1.0 / 8, np.float32))
    iters = 60
    x = x0.copy()
    for _ in range(iters):
        y = port.spmv_q824(ip2, indices, words, x)
        x = hsoracle.axpb_q824(alpha, y, beta)
    ctxs = [capi.Context(0, capi.IMPL_FIXED) for _ in range(2)]
    for c in ctxs:
        c.upload_matrix_csr(r2, c2, ip2, indices, words)
        c.upload_vector(x0)
        c.sync()
    errors = []

    def drive(c):
        try:
            for _ in range(6):
                c.iterate(iters // 6, alpha, beta)
            if not np.array_equal(c.download_result(), y):
                errors.append("mismatch")
        except Exception as e:                      # noqa: BLE001
            errors.append(repr(e))

    th = [threading.Thread(target=drive, args=(c,)) for c in ctxs]
    for t in th: t.start()
    for t in th: t.join()
    for c in ctxs: c.close()
    assert not errors, errors


@pytest.mark.parametrize("narrow", [0, 1], ids=["wide", "narrow"])
def test_iterate_peers_one_kernel_single_rank(gpu, port, monkeypatch, narrow):
    """hsb_iterate_peers with a world of one: the resident kernel stores its slice through the peer table, raises its own
    arrival flag from the last CTA of the second barrier and starts the next iteration on that flag; mixed with the
    launch-per-step calls (the arrival sequence numbers continue), bit-exact"""
    monkeypatch.setenv("HSB_NARROW", str(narrow))
    r2, c2, ip2, indices, data = _pagerank_matrix(6000, 90000, 73)
    words = port.quantize(data)
    alpha, beta = int(port.quantize(np.float32([0.85]))[0]), int(port.quantize(np.float32([0.02]))[0])
    x0 = port.quantize(np.full(c2, 0.125, np.float32))
    ctx = capi.Context(0, capi.IMPL_FIXED)
    ctx.upload_matrix_csr(r2, c2, ip2, indices, words)
    ctx.upload_vector(x0)
    ctx.peer_connect(1, 0, ctx.peer_export())
    x = x0.copy()

    def advance(n):
        nonlocal x
        for _ in range(n):
            y = port.spmv_q824(ip2, indices, words, x)
            x = hsoracle.axpb_q824(alpha, y, beta)
        return y

    ctx.iterate_peers(5, alpha, beta, 0)
    assert np.array_equal(ctx.download_result(), advance(5))
    for _ in range(2):                                    # launch-per-step form in between
        ctx.spmv(); ctx.axpb_to_peers(alpha, beta, 0); ctx.vector_commit()
    advance(2)
    ctx.iterate_peers(4, alpha, beta, 0)                  # starts on a vector announced by arrival flags (wait_first)
    assert np.array_equal(ctx.download_result(), advance(4))
    ctx.spmv()
    assert np.array_equal(ctx.download_result(), port.spmv_q824(ip2, indices, words, x))
    ctx.close()


def test_iterate_peers_one_kernel_two_gpus(gpu, port):
    """two GPUs, one host thread and one context each, the matrix split into two row blocks: every iteration of the
    resident kernels exchanges the slices over NVLink (peer stores + arrival flags). Skipped on a one-GPU box."""
    import threading
    from hisparse_b200 import sharding
    if capi.device_count() < 2:
        pytest.skip("needs two GPUs")
    r2, c2, ip2, indices, data = _pagerank_matrix(40000, 1200000, 79)
    words = port.quantize(data)
    alpha, beta = int(port.quantize(np.float32([0.85]))[0]), int(port.quantize(np.float32([0.02]))[0])
    x0 = port.quantize(np.full(c2, 0.125, np.float32))
    iters = 40
    x = x0.copy()
    for _ in range(iters):
        y = port.spmv_q824(ip2, indices, words, x)
        x = hsoracle.axpb_q824(alpha, y, beta)
    want_next = port.spmv_q824(ip2, indices, words, x)
    bounds = sharding.shard_bounds(ip2, 2)
    ctxs = []
    for g in range(2):
        sip, six, sw = sharding.extract_shard(ip2, indices, words, bounds[g], bounds[g + 1])
        c = capi.Context(g, capi.IMPL_FIXED)
        c.upload_matrix_csr(bounds[g + 1] - bounds[g], c2, sip, six, sw)
        c.upload_vector(x0)
        c.sync()
        ctxs.append(c)
    blobs = np.concatenate([c.peer_export() for c in ctxs])
    for g, c in enumerate(ctxs):
        c.peer_connect(2, g, blobs)
    errors = []

    def drive(g):
        try:
            c = ctxs[g]
            c.iterate_peers(iters // 2, alpha, beta, bounds[g])
            c.iterate_peers(iters - iters // 2, alpha, beta, bounds[g])
            if not np.array_equal(c.download_result(), y[bounds[g]:bounds[g + 1]]):
                errors.append((g, "y of the last iteration"))
            c.spmv()
            if not np.array_equal(c.download_result(), want_next[bounds[g]:bounds[g + 1]]):
                errors.append((g, "SpMV on the iterated vector"))
        except Exception as e:                      # noqa: BLE001
            errors.append((g, repr(e)))

    th = [threading.Thread(target=drive, args=(g,)) for g in range(2)]
    for t in th: t.start()
    for t in th: t.join()
    for c in ctxs: c.close()
    assert not errors, errors


def test_iterate_split_over_several_launches(gpu):
    """a long run is cut into several cooperative launches (HSB_ITERATE_CHUNK=3 here, 2^20 iterations normally): the
    vector, the buffer rotation and -- with peers -- the arrival sequence numbers carry over from one launch to the next.
    (A subprocess: the chunk size is read once per process.)"""
    import subprocess
    import sys
    code = r'''
import numpy as np, sys
sys.path.insert(0, %r)
from hisparse_b200 import capi, matgen
from oracle import hsoracle
port = hsoracle.Port()
rows, cols, indptr, indices, _ = matgen.rmat_csr(3000, 50000, 83)
outdeg = np.maximum(np.bincount(indices, minlength=cols), 1)
data = (1.0 / outdeg[indices]).astype(np.float32)
r2, c2, ip2 = matgen.pad_csr(rows, cols, indptr, 128, 128)
words = port.quantize(data)
alpha, beta = int(port.quantize(np.float32([0.85]))[0]), int(port.quantize(np.float32([0.02]))[0])
x0 = port.quantize(np.full(c2, 0.125, np.float32))
x = x0.copy()
for _ in range(10):
    y = port.spmv_q824(ip2, indices, words, x)
    x = hsoracle.axpb_q824(alpha, y, beta)
for peers in (False, True):
    ctx = capi.Context(0, "fixed")
    ctx.upload_matrix_csr(r2, c2, ip2, indices, words)
    ctx.upload_vector(x0)
    if peers:
        ctx.peer_connect(1, 0, ctx.peer_export())
        ctx.iterate_peers(10, alpha, beta, 0)
    else:
        ctx.iterate(10, alpha, beta)
    assert np.array_equal(ctx.download_result(), y), peers
    ctx.spmv()
    assert np.array_equal(ctx.download_result(), port.spmv_q824(ip2, indices, words, x)), peers
    assert ctx.stats()["kernel_launches"] >= 4
    ctx.close()
print("ok")
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, HSB_ITERATE_CHUNK="3")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]
