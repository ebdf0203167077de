"""A short walk through the product's GPU paths on small inputs, meant to run under compute-sanitizer
(memcheck / racecheck / synccheck / initcheck): GPU formatter, SpMV fixed and float, pipelined
upload -> SpMV -> download with host buffers, row partitions, CPSR channel-image ingestion, the iterative
caller, and the gather of y across two row-block contexts. Every result is checked against the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # repo root (this file lives in tests/: it uses the oracle as its checker)
sys.path.insert(0, ROOT)
from hisparse_b200 import capi, matgen, sharding  # noqa: E402
from oracle import hsoracle  # noqa: E402

port = hsoracle.Port()
rows, cols, indptr, indices, data = matgen.rmat_csr(3000, 60000, 5)
data = (data * np.float32(0.05)).astype(np.float32)
r2, c2, ip2 = matgen.pad_csr(rows, cols, indptr, 128, 8)
words = port.quantize(data)
rng = np.random.default_rng(1)
xs = [port.quantize(rng.random(c2, dtype=np.float32)) for _ in range(4)]
wants = [port.spmv_q824(ip2, indices, words, x) for x in xs]

# fixed: blocking sequence, then the pipelined one with page-locked buffers
ctx = capi.Context(0, "fixed")
ctx.upload_matrix_csr(r2, c2, ip2, indices, words, 1024)          # 3 row partitions
ctx.upload_vector(xs[0]); ctx.spmv()
assert np.array_equal(ctx.download_result(), wants[0])
px = [capi.PinnedArray(c2) for _ in range(2)]
py = [capi.PinnedArray(r2) for _ in range(4)]
for k in range(4):
    px[k & 1].array[:] = xs[k]
    ctx.upload_vector(px[k & 1].array); ctx.spmv(); ctx.download_result_async(py[k].array)
    if k & 1:
        ctx.sync()                                                # the host buffer is rewritten next round
ctx.sync()
for k in range(4):
    assert np.array_equal(py[k].array, wants[k]), k
for rp in range(3):
    n = min(1024, r2 - rp * 1024)
    ctx.spmv_row_partition(rp, n // 16, 1, 3, c2)
assert np.array_equal(ctx.download_result(), wants[3])
# iterative caller
alpha, beta = int(port.quantize(np.float32([0.85]))[0]), int(port.quantize(np.float32([0.02]))[0])
sq = matgen.rmat_csr(2048, 30000, 6)
sr2, sc2, sip2 = matgen.pad_csr(sq[0], sq[1], sq[2], 128, 8)
sw = port.quantize((sq[4] * np.float32(0.05)).astype(np.float32))
x = port.quantize(np.full(sc2, 0.125, np.float32))
ctx.upload_matrix_csr(sr2, sc2, sip2, sq[3], sw)
ctx.upload_vector(x)
ctx.iterate(3, alpha, beta)
for _ in range(3):
    x = hsoracle.axpb_q824(alpha, port.spmv_q824(sip2, sq[3], sw, x), beta)
ctx.spmv()
assert np.array_equal(ctx.download_result(), port.spmv_q824(sip2, sq[3], sw, x))
ctx.set_option("iterate_persistent", 0)                           # the launch-per-step form continues the iteration
ctx.iterate(2, alpha, beta)
for _ in range(2):
    x = hsoracle.axpb_q824(alpha, port.spmv_q824(sip2, sq[3], sw, x), beta)
ctx.spmv()
assert np.array_equal(ctx.download_result(), port.spmv_q824(sip2, sq[3], sw, x))
ctx.close()

# float
xf = (rng.random(c2, dtype=np.float32) * 2 - 1).astype(np.float32)
ctx = capi.Context(0, "float_pob")
ctx.upload_matrix_csr(r2, c2, ip2, indices, data)
ctx.upload_vector(xf); ctx.spmv()
y = ctx.download_result().view(np.float32).astype(np.float64)
y64, sa = port.spmv_f64(ip2, indices, data, xf)
assert np.all(np.abs(y - y64) <= 1e-5 * sa + 1e-30)
ctx.close()

# CPSR channel images (the reference's wire format) decoded on the device
IF, OB, VB = 1, 1 << 20, 32768
m = port.csr2cpsr(r2, c2, ip2, indices, words, 8, OB, VB, 16 * IF, True, hsoracle.VAL_Q824)
ctx = capi.Context(0, "fixed")
ctx.upload_matrix_cpsr(m.channel_images(IF), m.n_row_parts, m.n_col_parts, r2, c2)
ctx.upload_vector(xs[1]); ctx.spmv()
assert np.array_equal(ctx.download_result(), wants[1])
ctx.close()

# two row-block contexts, y gathered by the drains
bounds = sharding.shard_bounds(ip2, 2)
ctxs = []
for g in range(2):
    sip, six, sw_ = sharding.extract_shard(ip2, indices, words, bounds[g], bounds[g + 1])
    c = capi.Context(0, "fixed")
    c.upload_matrix_csr(bounds[g + 1] - bounds[g], c2, sip, six, sw_)
    c.upload_vector(xs[2])
    ctxs.append(c)
blobs = np.concatenate([c.gather_export(r2, want_buffer=(g == 0)) for g, c in enumerate(ctxs)])
for g, c in enumerate(ctxs):
    c.gather_connect(2, g, bounds[g], blobs)
for c in ctxs:
    c.spmv(); c.spmv()
for c in ctxs:
    c.sync()
assert np.array_equal(ctxs[0].download_gathered(), wants[2])
for c in ctxs:
    c.close()
print("sanitize_run: all paths agree with the oracle")
