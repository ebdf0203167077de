"""Device-resident SpMV time of the bench matrix (C2) for one library build (HSB_LIB) and one setting of the
run-time knobs (HSB_SLICE_COST, ...). The matrix is cached in /tmp between invocations. Prints one line."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from hisparse_b200 import capi, matgen  # noqa: E402

WL = os.environ.get("SWEEP_WORKLOAD", "c2")
bench.WORKLOAD = WL
cache = "/tmp/hsb_%s_cache.npz" % WL
if os.path.exists(cache):
    d = np.load(cache)
    r2, c2, ip2, indices, words, xw = int(d["r2"]), int(d["c2"]), d["ip2"], d["indices"], d["words"], d["xw"]
else:
    r2, c2, ip2, indices, data, x = bench.workload(0)
    if bench.WORKLOADS[WL][1] == "fixed":
        words, xw = matgen.quantize_q824(data), matgen.quantize_q824(x)
    else:
        words, xw = data.view(np.uint32), x.view(np.uint32)
    np.savez(cache, r2=r2, c2=c2, ip2=ip2, indices=indices, words=words, xw=xw)
ctx = capi.Context(0, bench.WORKLOADS[WL][1])
ctx.upload_matrix_csr(r2, c2, ip2, indices, words)
ctx.set_replicas(max(2, int(np.ceil(2.5 * (capi.device_l2_bytes(0) or 126 * 2 ** 20) / max(ctx.stats()["format_bytes"], 1)))))
ctx.upload_vector(xw)
n = 2048 if WL in ("c1", "c2", "c3") else 400
ts = [ctx.time_spmv(n // 8, n, kernel=False)[0] * 1e3 for _ in range(3)]
st = ctx.stats()
print("%s tiles=%d streams=%d fmtB/nnz=%.2f" % (WL, st["n_col_tiles"], st["n_streams"], st["format_bytes"] / max(st["nnz"], 1)), end="  ")
print("%-28s %-24s us/spmv %s  best %.2f" % (os.path.basename(os.environ.get("HSB_LIB", "default")),
                                            " ".join("%s=%s" % (k, v) for k, v in os.environ.items() if k.startswith("HSB_") and k != "HSB_LIB"),
                                            " ".join("%.2f" % t for t in ts), min(ts)))
ctx.close()
