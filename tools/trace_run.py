"""Per-warp / per-CTA finish times of one SpMV launch (SM clocks): python tools/trace_run.py --config c2"""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hisparse_b200 import capi, matgen  # noqa
from tools.profile_run import make  # noqa

ap = argparse.ArgumentParser(); ap.add_argument("--config", default="c2"); ap.add_argument("--impl", default="fixed")
a = ap.parse_args()
rows, cols, indptr, indices, data = make(a.config)
r2, c2, ip2 = matgen.pad_csr(rows, cols, indptr, 128, 8)
x = np.zeros(c2, np.float32); x[:cols] = np.random.default_rng(1).random(cols, dtype=np.float32)
if a.impl == "fixed":
    data, x = matgen.quantize_q824(data * np.float32(0.05)), matgen.quantize_q824(x)
ctx = capi.Context(0, a.impl)
ctx.upload_matrix_csr(r2, c2, ip2, indices, data)
ctx.set_replicas(4); ctx.upload_vector(x)
for _ in range(5): ctx.spmv()
ctx.sync(); ctx.trace(True)
for _ in range(3): ctx.spmv()
t = ctx.trace().astype(np.float64)
nw = t.shape[1] - 2
w = t[:, :nw]; arrive = t[:, nw]; done = t[:, nw + 1]
steps, slices = ctx.plan()
print("per-CTA: steps slices | warp finish min / mean / max, CTA done  (SM cycles; run with HSB_NO_PDL=1)")
for b in range(t.shape[0]):
    ww = w[b][w[b] > 0]
    print(b, steps[b], slices[b], "|", int(ww.min()) if ww.size else 0, int(ww.mean()) if ww.size else 0, int(ww.max()) if ww.size else 0, int(arrive[b]))
np.set_printoptions(linewidth=260)
print("CTA done (cycles/100) by CTA:", (done / 100).astype(int))
for b in (0, 20, 36, 50, 100):
    print("CTA", b, "per-warp finish/100:", (w[b] / 100).astype(int))
print("mean over CTAs per warp index /100:", (w.mean(0) / 100).astype(int))
# least-squares fit  time = a + b*steps + c*slices  over CTAs (mean warp finish time)
tm = np.array([w[b][w[b] > 0].mean() for b in range(t.shape[0])])
A = np.stack([np.ones_like(tm), steps.astype(np.float64), slices.astype(np.float64)], 1)
coef, *_ = np.linalg.lstsq(A, tm, rcond=None)
print("fit mean-warp time = %.0f + %.2f*steps + %.2f*slices  -> slice cost = %.2f steps; residual rms %.0f" % (coef[0], coef[1], coef[2], coef[2] / coef[1], np.sqrt(np.mean((A @ coef - tm) ** 2))))
print("arrive: min %d mean %d max %d ; done max %d" % (arrive.min(), arrive.mean(), arrive.max(), done.max()))
