"""Where does the host-buffer pipeline spend its time? Each leg of upload / SpMV / download alone and in
combination, flag pipeline on and off (hsb_set_option), on the bench matrix (C2). Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from hisparse_b200 import capi, matgen  # noqa: E402


def main():
    r2, c2, ip2, indices, data, x = bench.workload(0)
    words, xw = matgen.quantize_q824(data), matgen.quantize_q824(x)
    ctx = capi.Context(0, "fixed")
    ctx.upload_matrix_csr(r2, c2, ip2, indices, words)
    ctx.set_replicas(4)
    px = [capi.PinnedArray(c2) for _ in range(2)]
    py = [capi.PinnedArray(r2) for _ in range(2)]
    for b in px:
        b.array[:] = xw
    xh, yh = [b.array for b in px], [b.array for b in py]
    out = {}

    def timed(fn, n):
        ctx.sync()
        t0 = time.perf_counter()
        for k in range(n):
            fn(k)
        ctx.sync()
        return 1e6 * (time.perf_counter() - t0) / n

    for flags, once in ((1, 1), (1, 0), (0, 1)):          # second field: host_drain
        try:
            ctx.set_option("flags", flags)
            ctx.set_option("host_drain", once)
        except capi.HsbError as e:
            out["flags=%d" % flags] = str(e)
            continue
        ctx.upload_vector(xh[0]); ctx.spmv(); ctx.sync()
        r = {}
        r["resident_us"] = 1e3 * ctx.time_spmv(64, 1024, kernel=False)[0]
        r["upload_only_us"] = timed(lambda k: ctx.upload_vector(xh[k & 1]), 512)
        ctx.spmv(); ctx.sync()
        r["download_only_us"] = timed(lambda k: ctx.download_result_async(yh[k & 1]), 512)

        def up_spmv(k):
            ctx.upload_vector(xh[k & 1]); ctx.spmv()
        r["upload_spmv_us"] = timed(up_spmv, 1024)

        def spmv_down(k):
            ctx.spmv(); ctx.download_result_async(yh[k & 1])
        r["spmv_download_us"] = timed(spmv_down, 1024)
        r["full_async_us"] = 1e6 * ctx.time_e2e(xh, yh, 2048, async_download=True)
        r["full_sync_us"] = 1e6 * ctx.time_e2e(xh, yh, 256, async_download=False)
        out["flags=%d,host_drain=%d" % (flags, once)] = r
    print(json.dumps(out))
    ctx.close()


if __name__ == "__main__":
    main()
