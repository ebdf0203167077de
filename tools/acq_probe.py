"""Cost of the acquire semantics of the kernels' flag waits: device-resident loop and host-buffer pipeline on the
bench matrix (C2) with hsb_set_option("acquire", 1 | 0), for the library HSB_LIB selects (build with
-DHSB_ACQ_FENCE=1 for the relaxed-load-plus-fence variant). Prints one line per setting."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from hisparse_b200 import capi, matgen  # noqa: E402

cache = "/tmp/hsb_c2_cache.npz"
if os.path.exists(cache):
    d = np.load(cache)
    r2, c2, ip2, indices, words, xw = int(d["r2"]), int(d["c2"]), d["ip2"], d["indices"], d["words"], d["xw"]
else:
    r2, c2, ip2, indices, data, x = bench.workload(0)
    words, xw = matgen.quantize_q824(data), matgen.quantize_q824(x)
    np.savez(cache, r2=r2, c2=c2, ip2=ip2, indices=indices, words=words, xw=xw)
ctx = capi.Context(0, "fixed")
ctx.upload_matrix_csr(r2, c2, ip2, indices, words)
ctx.set_replicas(4)
px = [capi.PinnedArray(c2) for _ in range(2)]
py = [capi.PinnedArray(r2) for _ in range(2)]
for b in px:
    b.array[:] = xw
xh, yh = [b.array for b in px], [b.array for b in py]
for acq, xfc in ((1, 1), (1, 0), (1, 1), (1, 0), (0, 1)):
    ctx.set_option("acquire", acq)
    ctx.set_option("xflag_copy", xfc)
    ctx.upload_vector(xw)
    res = min(ctx.time_spmv(256, 2048, kernel=False)[0] for _ in range(3)) * 1e3
    e2e = min(ctx.time_e2e(xh, yh, 2048, async_download=True) for _ in range(3)) * 1e6
    sync = ctx.time_e2e(xh, yh, 256, async_download=False) * 1e6
    t0 = __import__("time").perf_counter()
    for k in range(512):
        ctx.upload_vector(xh[k & 1])
    ctx.sync()
    up = (__import__("time").perf_counter() - t0) / 512 * 1e6
    print("%-24s acquire=%d xflag_copy=%d  resident %.2f us  upload only %.2f us  e2e pipelined %.2f us  synchronous %.2f us"
          % (os.path.basename(os.environ.get("HSB_LIB", "default")), acq, xfc, res, up, e2e, sync), flush=True)
ctx.close()
