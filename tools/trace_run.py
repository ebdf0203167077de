"""Per-warp / per-CTA finish times of one SpMV launch (SM clocks): python tools/trace_run.py --config c2"""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hisparse_b200 import capi, matgen  # noqa
from oracle import hsoracle  # noqa
from tools.profile_run import make  # noqa

ap = argparse.ArgumentParser(); ap.add_argument("--config", default="c2"); ap.add_argument("--impl", default="fixed")
a = ap.parse_args()
rows, cols, indptr, indices, data = make(a.config)
r2, c2, ip2 = matgen.pad_csr(rows, cols, indptr, 128, 8)
x = np.zeros(c2, np.float32); x[:cols] = np.random.default_rng(1).random(cols, dtype=np.float32)
if a.impl == "fixed":
    port = hsoracle.Port(); data, x = port.quantize(data * np.float32(0.05)), port.quantize(x)
ctx = capi.Context(0, a.impl)
ctx.upload_matrix_csr(r2, c2, ip2, indices, data)
ctx.set_replicas(4); ctx.upload_vector(x)
for _ in range(5): ctx.spmv()
ctx.sync(); ctx.trace(True)
for _ in range(3): ctx.spmv()
t = ctx.trace().astype(np.float64)
w = t[:, :32]; arrive = t[:, 32]; done = t[:, 33]
print("per-CTA: warp finish min / mean / max, barrier arrive, drain done  (SM cycles)")
for b in range(t.shape[0]):
    ww = w[b][w[b] > 0]
    print(b, int(ww.min()) if ww.size else 0, int(ww.mean()) if ww.size else 0, int(ww.max()) if ww.size else 0, int(arrive[b]), int(done[b]))
print("arrive: min %d mean %d max %d ; done max %d" % (arrive.min(), arrive.mean(), arrive.max(), done.max()))
