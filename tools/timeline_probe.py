"""Launch-by-launch timeline (globaltimer) of the SpMV kernel in three regimes: device-resident loop,
upload + SpMV, SpMV + deferred download, full pipeline. Prints medians of the per-launch intervals."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from hisparse_b200 import capi, matgen  # noqa: E402


def summarize(ctx, n):
    last, t = ctx.timeline()
    seqs = [s for s in range(last - n + 8, last - 2)]
    rows = np.array([t[s & 255] for s in seqs]).astype(np.float64)
    start, flag, prev, drain, xst, end, work = [rows[:, i] for i in range(7)]
    med = lambda a: float(np.median(a)) / 1e3
    return {"period_us": med(np.diff(end)), "duration_us": med(end - start),
            "start_to_flag_us": med(flag - start), "start_to_prev_done_us": med(prev - start),
            "prev_done_to_drain_us": med(drain - prev), "start_to_x_staged_us": med(xst - start),
            "prev_end_to_start_us": med(start[1:] - end[:-1]), "x_staged_after_prev_done_us": med(xst - prev),
            "cta0_start_to_work_done_us": med(work - start), "cta0_work_done_to_prev_done_us": med(prev - work),
            "cta0_prev_done_to_published_us": med(drain - prev)}


def main():
    wl = os.environ.get("TL_WORKLOAD", "c2")
    bench.WORKLOAD = wl
    r2, c2, ip2, indices, data, x = bench.workload(0)
    if os.environ.get("TL_SHARD"):                       # "world:rank": one nnz-balanced row block of the matrix
        from hisparse_b200 import sharding
        world, rank = [int(v) for v in os.environ["TL_SHARD"].split(":")]
        bounds = sharding.shard_bounds(ip2, world)
        ip2, indices, data = sharding.extract_shard(ip2, indices, data, bounds[rank], bounds[rank + 1])
        r2 = bounds[rank + 1] - bounds[rank]
    if bench.WORKLOADS[wl][1] == "fixed":
        words, xw = matgen.quantize_q824(data), matgen.quantize_q824(x)
    else:
        words, xw = data.view(np.uint32), x.view(np.uint32)
    ctx = capi.Context(0, bench.WORKLOADS[wl][1])
    ctx.upload_matrix_csr(r2, c2, ip2, indices, words)
    ctx.set_replicas(max(2, int(np.ceil(2.5 * (capi.device_l2_bytes(0) or 126 * 2 ** 20) / max(ctx.stats()["format_bytes"], 1)))))
    px = [capi.PinnedArray(c2) for _ in range(2)]
    py = [capi.PinnedArray(r2) for _ in range(2)]
    for b in px:
        b.array[:] = xw
    xh, yh = [b.array for b in px], [b.array for b in py]
    out = {}
    n = 200
    ctx.upload_vector(xh[0]); ctx.spmv(); ctx.sync()
    ctx.timeline(True)
    ctx.time_spmv(0, n, kernel=False)
    out["resident"] = summarize(ctx, n)
    ctx.timeline(True)
    for k in range(n):
        ctx.upload_vector(xh[k & 1]); ctx.spmv()
    ctx.sync()
    out["upload_spmv"] = summarize(ctx, n)
    ctx.timeline(True)
    for k in range(n):
        ctx.spmv(); ctx.download_result_async(yh[k & 1])
    ctx.sync()
    out["spmv_download"] = summarize(ctx, n)
    ctx.timeline(True)
    ctx.time_e2e(xh, yh, n, async_download=True)
    out["full"] = summarize(ctx, n)
    print(json.dumps(out, indent=1))
    ctx.close()


if __name__ == "__main__":
    main()
