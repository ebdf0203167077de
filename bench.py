#!/usr/bin/env python
"""bench.py -- SpMV throughput of the B200 hot path on BASELINE.json's metric and config.

    python bench.py --gpus N --steps K --warmup W            (ours; N>1 under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (N=1): BASELINE.json configs[1] -- googleplus-sized graph (107,614 x 107,614, ~13.7 M
non-zeros; the dataset itself is a download the reference does not ship, so an R-MAT stand-in of
the same size is generated, seed 0xC0FFEE02), fixed-point path (ap_ufixed<32,8>), one B200.
A STEP is one batch of `--batch` SpMVs (default 512) over the resident matrix -- the reference's
benchmark loop (sw/benchmark.cpp:315-343) with 512 instead of 50 back-to-back runs -- so that a
step lasts milliseconds and GPU clocks can be sampled while it runs. Successive SpMVs rotate over
enough HBM copies of the matrix that none is still in the L2 (126 MB on B200, queried) when it is read again.

N > 1 (torchrun): `value` stays the same workload, one C2-sized matrix per GPU (so that the N=1 line is the
first point of the curve), and a `sharded` object is added -- BASELINE.json's multi-GPU configurations run the
way north_star splits them: ONE matrix cut into nnz-balanced row blocks (C4, ogbl-ppa-sized, at 2 and 4 GPUs;
the device-generated C5, 100 M x 100 M power-law, at 8), x broadcast with NCCL, y gathered ON THE DEVICE by the
result drain's peer stores over NVLink, the gathered y checked against the oracle. See sharded_block().

Metric (sw/benchmark.cpp:311-346): GOPS = 2*nnz / t ; GBPS = 8*nnz bytes / 2^30 / t.
`value` = device-timed (CUDA events on the launching stream, inputs resident in HBM);
`e2e`   = the same metric through the reference-facing C ABI with HOST buffers: every SpMV uploads x
          from pinned host memory and downloads y (matrix resident, as in the reference).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 0xC0FFEE02
N_NODES, NNZ_TARGET = 107614, 13_670_000
L2_BYTES = 126 * 1024 * 1024     # replaced by the device's own figure (hsb_device_l2_bytes) once a GPU is open


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


WORKLOADS = {
    # name: (description, impl, generator)            -- SURVEY.md section 8d; c2 is the bench line
    "c1": ("C1: 4096x4096 uniform random, 1 %% density, fp32", "float_pob",
           lambda m, r: m.random_csr(4096, 4096, 0.01, 0xC0FFEE01 + r)),
    "c2": ("C2: googleplus-sized R-MAT 107614^2 (a,b,c=.57,.19,.19), fixed-point", "fixed",
           lambda m, r: m.rmat_csr(N_NODES, NNZ_TARGET, SEED + r)),
    "c3": ("C3: transformer-sized 512x33288 Bernoulli mask 50 %%, fp32 (float_pob)", "float_pob",
           lambda m, r: m.bernoulli_csr(512, 33288, 0.5, 0xC0FFEE03 + r)),
    "t95": ("transformer-sized 512x33288 Bernoulli mask 5 %%, fp32 (float_pob)", "float_pob",
            lambda m, r: m.bernoulli_csr(512, 33288, 0.05, 0xC0FFEE03 + r)),
    # C5: one row-block shard per GPU of the 100 M x 100 M power-law matrix (12.5 M rows, ~245 M non-zeros each: the
    # whole matrix at 8 GPUs), generated AND formatted on the device (hsb_synth_powerlaw_csr_device)
    "c5s": ("C5 shard: rows [rank*12.5M, +12.5M) of the 100M x 100M power-law matrix (alpha 2.1, mean degree 20, "
            "80 %% of the columns within +-2^20 of the diagonal), generated on the device, fp32", "float_pob", None),
    "c4": ("C4: ogbl-ppa-sized symmetric R-MAT 576289^2, fp32", "float_pob",
           lambda m, r: m.rmat_csr(576289, 42_460_000, 0xC0FFEE04 + r, symmetric=True, oversample=1.5)),
}
WORKLOAD = "c2"


SHARD_ONE_MATRIX = False


def workload_name(nnz):
    """config.workload: the same string in both arms (ours / --impl reference)"""
    return (WORKLOADS[WORKLOAD][0] % ()) + ", nnz=%d" % nnz


def workload(rank, world=1):
    """Synthetic input of `rank`. Default (weak scaling): every rank gets its own matrix of the workload's
    size (rank 0 == the single-GPU workload). --shard-one-matrix (strong scaling, BASELINE config C4):
    ONE matrix, cut into nnz-balanced row blocks (hisparse_b200/sharding.py), rank g keeps block g."""
    from hisparse_b200 import matgen, sharding
    rows, cols, indptr, indices, data = WORKLOADS[WORKLOAD][2](matgen, 0 if SHARD_ONE_MATRIX else rank)
    if WORKLOADS[WORKLOAD][1] == "fixed":
        data = (data * np.float32(0.05)).astype(np.float32)  # keeps most row sums below saturation
    r2, c2, ip2 = matgen.pad_csr(rows, cols, indptr, 128, 8)  # util_round_csr_matrix_dim
    x = np.zeros(c2, np.float32)
    x[:cols] = np.random.default_rng(SEED).random(cols, dtype=np.float32)
    if SHARD_ONE_MATRIX and world > 1:
        bounds = sharding.shard_bounds(ip2, world)
        ip2, indices, data = sharding.extract_shard(ip2, indices, data, bounds[rank], bounds[rank + 1])
        r2 = bounds[rank + 1] - bounds[rank]
    return r2, c2, ip2, indices, data, x


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        threading.Thread(target=self._pump, daemon=True).start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        inside = [r for t, r in self.rows if self.t0 <= t <= self.t1]
        scope = "timed region"
        if not inside:
            inside, scope = [r for _, r in self.rows], "whole run (timed region shorter than the sampling period)"
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in inside:
            f = [s.strip() for s in r.split(",")]
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "scope": scope}


def _host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def _reference_timers(r2, c2, ip2, indices, data, x):
    """-> (kind, one(n_runs, threads) -> seconds per SpMV). The reference's CPU SpMV is compute_ref
    (sw/host.cpp:33-48), a single-thread loop; `threads` > 1 runs that unmodified function on
    nnz-balanced row blocks, one block per host thread (oracle/ref_driver.cpp)."""
    from oracle import hsoracle
    if hsoracle.ref_available("fixed"):
        ref = hsoracle.Ref("fixed")

        def one(n, threads):
            if threads <= 1:
                return ref.time_compute_ref(r2, c2, ip2, indices, data, x, n)
            return ref.time_compute_ref_mt(r2, c2, ip2, indices, data, x, n, threads)
        return "reference", one
    hsoracle.build()
    port = hsoracle.Port()
    return "port", (lambda n, threads: port.time_spmv_f32(ip2, indices, data, x, n))


def run_reference(args, emit):
    """The reference's own CPU implementation of the path, on all the host threads it can use."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if WORKLOAD == "c5s":
        raise SystemExit("bench: the C5 shard is generated on the GPU; the reference arm runs the host-generated workloads")
    r2, c2, ip2, indices, data, x = workload(0)
    nnz = int(ip2[-1])
    kind, one = _reference_timers(r2, c2, ip2, indices, data, x)
    threads = _host_threads() if kind == "reference" else 1
    t_single = one(2, 1)
    # one thread per core is not always the fastest way to run a gather-bound loop: keep the better of
    # "all threads" and "as shipped" (1 thread) as the reference's figure
    if threads > 1 and one(2, threads) > t_single:
        threads = 1
    # a step = the same batch of SpMVs as our arm when that stays within ~8 s of CPU time, else a bounded sample
    t_one = max(one(2, threads), 1e-6)
    per_step = args.batch if args.batch * t_one <= 8.0 else max(4, int(2.0 / t_one))
    for _ in range(args.warmup):
        one(1, threads)
    t = 0.0
    for _ in range(args.steps):
        t += one(per_step, threads) * per_step
    sec_per_spmv = t / (args.steps * per_step)
    gops = 2.0 * nnz / sec_per_spmv / 1e9
    out = {"impl": "reference", "metric": "SpMV GOPS (2*nnz/t, sw/benchmark.cpp:312-346)", "value": gops,
           "unit": "GOPS", "gbps": 8.0 * nnz / 2 ** 30 / sec_per_spmv, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": workload_name(nnz), "spmv_per_step": per_step,
                      "arm": "fp32 compute_ref (sw/host.cpp:33-48) on %d host thread(s)" % threads},
           "cpu_baseline": {"value": gops, "unit": "GOPS", "cores": threads, "kind": kind,
                            "single_thread_value": 2.0 * nnz / t_single / 1e9,
                            "sample": "%d x %d full SpMVs of the matrix" % (args.steps, per_step)},
           "e2e": {"value": gops, "unit": "GOPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


def cpu_baseline(r2, c2, ip2, indices, data, x, nnz):
    kind, one = _reference_timers(r2, c2, ip2, indices, data, x)
    threads = _host_threads() if kind == "reference" else 1
    t1 = one(2, 1)
    runs1 = int(max(4, min(1000, 4.0 / max(t1, 1e-6))))        # about 4 s single thread + 8 s all threads
    sec1 = one(runs1, 1)
    sec, used = sec1, 1
    if threads > 1:
        tm = one(2, threads)
        secm = one(int(max(4, min(4000, 8.0 / max(tm, 1e-6)))), threads)
        if secm < sec1:
            sec, used = secm, threads
    return {"value": 2.0 * nnz / sec / 1e9, "unit": "GOPS", "gbps": 8.0 * nnz / 2 ** 30 / sec, "cores": used,
            "kind": kind, "host_cores_available": threads, "single_thread_value": 2.0 * nnz / sec1 / 1e9,
            "sample": "full fp32 SpMVs of the bench matrix by compute_ref (sw/host.cpp:33-48): %d runs on 1 thread, "
                      "then on %d threads (nnz-balanced row blocks); the faster is reported" % (runs1, threads)}


class _DevArray:
    """expose a raw device pointer to torch through __cuda_array_interface__"""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<u4", "data": (ptr, False), "version": 2}


def pin_to_gpu_node(capi, device):
    """Run this rank's host thread on the CPUs of the NUMA node its GPU hangs off (sysfs), so that the per-SpMV
    uploads and downloads of N ranks do not all cross one socket; the page-locked buffers themselves are already
    allocated on that node (hsb_host_alloc). Silently keeps the inherited affinity when the platform does not say."""
    node = capi.lib().hsb_device_numa_node(device)
    try:
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return node, len(cpus)
    except (OSError, ValueError):
        return node, 0


def _max_over_ranks(dist, dev, *vals):
    import torch
    t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def _sum_over_ranks(dist, dev, *vals):
    import torch
    t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(v) for v in t]


def sharded_block(args, dist, rank, local, world):
    """BASELINE.json configs[3] / configs[4] the way north_star splits them: ONE matrix, nnz-balanced contiguous
    row blocks (hisparse_b200/sharding.py), rank g holds block g and a replica of x and produces y[block g].
    Measured, max over ranks, device timed:
      spmv_ms            shards resident, x resident, no exchange                       (a)
      bcast_x_ms         one ncclBroadcast of x from rank 0, alone                      (b)
      spmv_gather_ms     as (a) with y gathered ON THE DEVICE: every rank's result drain also stores its block
                         into rank 0's gathered vector (peer stores over NVLink) and raises an arrival flag   (c)
      iter_ms            ncclBroadcast(x) -> SpMV -> gather, every iteration, stream ordered (no host sync)
      parity             the gathered y of rank 0 against the oracle                    (d)
      strong_scaling_efficiency = t(1 GPU, whole matrix) / (N * spmv_ms)   (C4; rank 0 times the whole matrix)  (e)
    HSB_BENCH_SHARDED=c4|c5|off overrides the choice (default: c5 at 8 GPUs, c4 below)."""
    import torch
    from hisparse_b200 import capi, matgen, sharding
    from oracle import hsoracle
    which = os.environ.get("HSB_BENCH_SHARDED") or ("c5" if world >= 8 else "c4")
    if which == "off":
        return None
    dev = torch.device("cuda", local)
    port = hsoracle.Port()
    impl = "float_pob"
    out = {"n_gpus": world, "impl": "fp32 (float_pob semantics: multiply then add, not fused)"}
    t_setup = time.perf_counter()
    ctx = capi.Context(local, impl)
    one_gpu_ms = None
    if which == "c4":
        rows, cols, indptr, indices, data = matgen.rmat_csr(576289, 42_460_000, 0xC0FFEE04, symmetric=True, oversample=1.5)
        r_all, c2, ip_all = matgen.pad_csr(rows, cols, indptr, 128, 8)
        x = np.zeros(c2, np.float32)
        x[:cols] = np.random.default_rng(SEED).random(cols, dtype=np.float32)
        bounds = sharding.shard_bounds(ip_all, world)
        sip, six, sdata = sharding.extract_shard(ip_all, indices, data, bounds[rank], bounds[rank + 1])
        if rank == 0:
            # the single-GPU time of the SAME matrix, for the strong-scaling figure
            ctx.upload_matrix_csr(r_all, c2, ip_all, indices, data.view(np.uint32))
            ctx.set_replicas(max(2, int(np.ceil(2.5 * L2_BYTES / max(ctx.stats()["format_bytes"], 1)))))
            ctx.upload_vector(x.view(np.uint32))
            ctx.time_spmv(args.warmup * 64, 1, kernel=False)
            one_gpu_ms, _ = ctx.time_spmv(0, max(64, args.steps * 32), kernel=False)
        ctx.upload_matrix_csr(bounds[rank + 1] - bounds[rank], c2, sip, six, sdata.view(np.uint32))
        out["workload"] = ("C4: ogbl-ppa-sized symmetric R-MAT %d^2, nnz=%d, fp32, cut into %d nnz-balanced row blocks"
                           % (r_all, int(ip_all[-1]), world))
        nnz_total = int(ip_all[-1])
        windows = None
        B = 128
    else:
        r_shard = int(os.environ.get("HSB_BENCH_C5_ROWS", "12500000"))
        c2 = 100_000_000
        r_all = r_shard * world
        bounds = [g * r_shard for g in range(world + 1)]
        dcsr = capi.DeviceCsr.powerlaw(local, r_shard, c2, first_global_row=rank * r_shard, seed=0xC0FFEE05)
        ctx.upload_matrix_csr_device(dcsr)
        # parity windows: the first and the last 2^19 rows of every shard come to the host, the rest never does
        w = min(1 << 19, r_shard)
        windows = [(0, w) + dcsr.download_rows(0, w), (r_shard - w, r_shard) + dcsr.download_rows(r_shard - w, r_shard)]
        nnz_shard = int(dcsr.nnz)
        dcsr.free()
        x = np.random.default_rng(SEED).random(c2, dtype=np.float32)
        nnz_total = int(_sum_over_ranks(dist, dev, float(nnz_shard))[0])
        out["workload"] = ("C5: rows [0, %d) of the 100M-column power-law matrix (alpha 2.1, mean degree 20, 80 %% of a row's "
                           "columns within +-2^20 of the diagonal), nnz=%d, fp32, generated and formatted on the devices, "
                           "%d row blocks of %d rows" % (r_all, nnz_total, world, r_shard))
        B = 8
    st = ctx.stats()
    replicas = max(1, int(np.ceil(2.5 * L2_BYTES / max(st["format_bytes"], 1)))) if st["format_bytes"] < 4 * L2_BYTES else 1
    ctx.set_replicas(replicas)
    # x: on rank 0 only, then one NCCL broadcast straight into the engine's device buffer
    xw = x.view(np.uint32)
    ctx.upload_vector(xw if rank == 0 else np.zeros(c2, np.uint32))
    ctx.sync()
    xt = torch.as_tensor(_DevArray(ctx.device_x(), c2), device=dev).view(torch.int32)
    dist.broadcast(xt, 0)
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t_setup

    def barrier():
        ctx.sync()
        dist.barrier()
        torch.cuda.synchronize()

    n_timed = max(B, args.steps * B // 4)
    # (a) resident shards, no exchange
    ctx.time_spmv(args.warmup * B, 1, kernel=False)
    barrier()
    spmv_ms, _ = ctx.time_spmv(0, n_timed, kernel=False)
    barrier()
    # (b) the broadcast alone
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        dist.broadcast(xt, 0)
    torch.cuda.synchronize()
    n_b = 20
    e0.record()
    for _ in range(n_b):
        dist.broadcast(xt, 0)
    e1.record()
    torch.cuda.synchronize()
    bcast_ms = e0.elapsed_time(e1) / n_b
    # gather of y: the drains store into rank 0's gathered vector from now on
    blob = torch.from_numpy(ctx.gather_export(r_all, want_buffer=(rank == 0))).to(dev)
    blobs = [torch.empty_like(blob) for _ in range(world)]
    dist.all_gather(blobs, blob)
    ctx.gather_connect(world, rank, bounds[rank], np.concatenate([b.cpu().numpy() for b in blobs]))
    barrier()
    # (d) parity of the gathered vector
    ctx.spmv()
    barrier()                                              # every rank's drain (and its peer stores) has completed
    y_block = ctx.download_result()
    import zlib
    ok, parity = 1.0, ""
    crc = float(zlib.crc32(y_block.tobytes()))
    crcs = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(crcs, torch.tensor([crc], dtype=torch.float64, device=dev))
    if windows is None:
        if rank == 0:
            yg = ctx.download_gathered().view(np.float32).astype(np.float64)
            y64, sa = port.spmv_f64(ip_all, indices, data, x)
            ok = float(np.all(np.abs(yg - y64) <= 1e-5 * sa + 1e-30))
            parity = ("gathered y (%d rows on rank 0, assembled by the drains' peer stores) within 1e-5 * sum|a_i x_i| of "
                      "the oracle's fp64 SpMV of the WHOLE matrix, all rows, checked in this run" % r_all)
    else:
        yb = y_block.view(np.float32).astype(np.float64)
        for (r0, r1, wip, wix, wv) in windows:
            y64, sa = port.spmv_f64(wip, wix, wv.view(np.float32), x)
            ok = min(ok, float(np.all(np.abs(yb[r0:r1] - y64) <= 1e-5 * sa + 1e-30)))
        if rank == 0:
            yg = ctx.download_gathered()
            for g in range(world):
                ok = min(ok, float(float(zlib.crc32(yg[bounds[g]:bounds[g + 1]].tobytes())) == float(crcs[g][0])))
            parity = ("every rank: first and last %d rows of its block within 1e-5 * sum|a_i x_i| of the oracle's fp64 "
                      "SpMV; rank 0: every block of the gathered y (%d rows) bit-identical (CRC-32) to the owning rank's "
                      "own y; checked in this run" % (windows[0][1], r_all))
    ok = -_max_over_ranks(dist, dev, -ok)[0]
    if ok < 1.0:
        raise SystemExit("bench: sharded SpMV result differs from the oracle -- refusing to report a number")
    barrier()
    # (c) the same loop with the gather fused into the drains
    launches0 = ctx.stats()["kernel_launches"]
    ctx.time_spmv(B, 1, kernel=False)
    barrier()
    launches1 = ctx.stats()["kernel_launches"]
    gather_ms, _ = ctx.time_spmv(0, n_timed, kernel=False)
    barrier()
    launches2 = ctx.stats()["kernel_launches"]
    # iteration with a fresh x every time: ncclBroadcast(x) -> SpMV (+ gather), stream ordered on the engine's stream
    ext = torch.cuda.ExternalStream(ctx.stream(), device=dev)
    n_it = max(8, n_timed // 4)
    with torch.cuda.stream(ext):
        for _ in range(3):
            dist.broadcast(xt, 0)
            ctx.spmv()
        barrier()
        e0.record(ext)
        for _ in range(n_it):
            dist.broadcast(xt, 0)
            ctx.spmv()
        ctx.sync()
        e1.record(ext)
    torch.cuda.synchronize()
    iter_ms = e0.elapsed_time(e1) / n_it
    barrier()
    spmv_ms, gather_ms, iter_ms, bcast_ms = _max_over_ranks(dist, dev, spmv_ms, gather_ms, iter_ms, bcast_ms)
    alg_sum, fmt_sum = _sum_over_ranks(dist, dev, float(st["algorithmic_bytes"]), float(st["format_bytes"]))
    peak, peak_src = load_peaks()
    gops = lambda ms: 2.0 * nnz_total / (ms / 1e3) / 1e9
    out.update({
        "rows": r_all, "cols": c2, "nnz": nnz_total, "row_block_bounds": bounds if world <= 16 else None,
        "spmv_ms": spmv_ms, "gops": gops(spmv_ms), "gbps": 8.0 * nnz_total / 2 ** 30 / (spmv_ms / 1e3),
        "spmv_gather_ms": gather_ms, "gops_with_gather": gops(gather_ms),
        "bcast_x_ms": bcast_ms, "bcast_x_gbs": c2 * 4 / (bcast_ms / 1e3) / 1e9,
        "iter_ms": iter_ms, "gops_bcast_every_spmv": gops(iter_ms),
        "roofline": {"bound": "hbm", "unit": "GB/s", "peak": peak, "peak_source": peak_src,
                     "achieved_per_gpu": alg_sum / world / (spmv_ms / 1e3) / 1e9,
                     "frac": alg_sum / world / (spmv_ms / 1e3) / 1e9 / peak,
                     "format_frac": fmt_sum / world / (spmv_ms / 1e3) / 1e9 / peak,
                     "note": "algorithmic bytes of a shard (mean over ranks; every shard counts the whole x once) / the "
                             "slowest rank's time per SpMV"},
        "gpu_launches_per_spmv_with_gather": (launches2 - launches1) / float(n_timed),
        "spmv_timed": n_timed, "l2_policy": "%d HBM replica(s) of the shard round-robin (%.0f MB each)" % (replicas, st["format_bytes"] / 1e6),
        "parity": parity, "setup_s": setup_s,
        "exchange": "x: ncclBroadcast from rank 0 (torch.distributed, NCCL over NVLink) into the engine's device buffer; "
                    "y: no collective call -- the result drain at the end of every SpMV stores the rank's block into rank 0's "
                    "gathered vector through peer pointers (CUDA IPC) and raises an arrival flag (hsb_gather_connect)",
        "limiting_collective": ("ncclBroadcast(x): %.3f ms against %.3f ms per SpMV; the gather of y costs %.3f ms per SpMV "
                                "(fused into the drain)" % (bcast_ms, spmv_ms, gather_ms - spmv_ms)),
    })
    if one_gpu_ms is not None or which == "c4":
        t1 = _max_over_ranks(dist, dev, one_gpu_ms or 0.0)[0]
        out["one_gpu_ms"] = t1
        out["strong_scaling_speedup"] = t1 / spmv_ms
        out["strong_scaling_efficiency"] = t1 / (world * spmv_ms)
        out["strong_scaling_efficiency_with_gather"] = t1 / (world * gather_ms)
    ctx.close()
    return out if rank == 0 else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=512, help="SpMVs per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS),
                    help="c2 (default) is the bench line; the others are extra measurements")
    ap.add_argument("--shard-one-matrix", action="store_true",
                    help="N > 1: row-block shards of ONE matrix (strong scaling) instead of one matrix per rank")
    args = ap.parse_args()
    global WORKLOAD, SHARD_ONE_MATRIX, L2_BYTES
    WORKLOAD = args.workload
    SHARD_ONE_MATRIX = args.shard_one_matrix
    # stdout carries exactly ONE line, the JSON: whatever libraries print there (NCCL's version banner, ...) goes
    # to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    if args.impl == "reference":
        return run_reference(args, emit)

    from hisparse_b200 import capi
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from hisparse_b200 import matgen
    placement = None
    if world > 1:
        placement = pin_to_gpu_node(capi, local)
    L2_BYTES = capi.device_l2_bytes(local) or L2_BYTES
    impl = WORKLOADS[WORKLOAD][1]
    ctx = capi.Context(local, impl)
    if WORKLOAD == "c5s":
        # the shard never exists on the host in CSR form before the run; it is downloaded afterwards for the checks
        r2, c2 = 12_500_000, 100_000_000
        dcsr = capi.DeviceCsr.powerlaw(local, r2, c2, first_global_row=rank * r2, seed=0xC0FFEE05)
        ctx.upload_matrix_csr_device(dcsr)
        ip2, indices, words = dcsr.download()
        dcsr.free()
        data = words.view(np.float32)
        x = np.random.default_rng(SEED).random(c2, dtype=np.float32)
        xw = x.view(np.uint32)
    else:
        r2, c2, ip2, indices, data, x = workload(rank, world)
        if impl == "fixed":
            words, xw = matgen.quantize_q824(data), matgen.quantize_q824(x)  # host-side float -> VAL_T conversion
        else:
            words, xw = data.view(np.uint32), x.view(np.uint32)
        ctx.upload_matrix_csr(r2, c2, ip2, indices, words)
    nnz = int(ip2[-1])
    st = ctx.stats()
    replicas = max(2, int(np.ceil(2.5 * L2_BYTES / max(st["format_bytes"], 1))))
    ctx.set_replicas(replicas)
    if world > 1:
        # every rank needs the same x: NCCL broadcast from rank 0 straight into the engine's x buffer
        import torch
        ctx.upload_vector(xw if rank == 0 else np.zeros_like(xw))
        ctx.sync()
        xt = torch.as_tensor(_DevArray(ctx.device_x(), c2), device="cuda:%d" % local)
        dist.broadcast(xt.view(torch.int32), 0)
        torch.cuda.synchronize()
    else:
        ctx.upload_vector(xw)
    ctx.spmv()
    y = ctx.download_result()            # checked bit-for-bit against the oracle in the cpu_baseline leg (N=1)

    B = args.batch if WORKLOAD != "c5s" else max(1, min(args.batch, 16))       # a C5-shard SpMV takes milliseconds
    launches0 = ctx.stats()["kernel_launches"]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ctx.time_spmv(args.warmup * B, 1, kernel=False)

    def barrier():
        ctx.sync()
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    barrier()
    launches1 = ctx.stats()["kernel_launches"]
    sampler.t0 = time.perf_counter()
    step_ms, _ = ctx.time_spmv(0, args.steps * B, kernel=False)       # CUDA events around K*B SpMVs
    barrier()
    sampler.t1 = time.perf_counter()
    launches2 = ctx.stats()["kernel_launches"]
    total_ms = step_ms * args.steps * B
    # One SpMV is ONE launch of spmv_tiles_kernel (the previous launch's row drain is fused into its
    # prologue), so the kernel's average launch duration over the timed region is total / launches.
    kernel_ms = total_ms / (launches2 - launches1)
    # the same kernel launched in isolation (an event pair around every launch defeats the
    # programmatic-dependent-launch overlap between consecutive launches)
    _, kernel_ms_isolated = ctx.time_spmv(0, min(args.steps * B, 512 if WORKLOAD != "c5s" else 8), kernel=True)
    if dist is not None:
        import torch
        t = torch.tensor([total_ms, float(nnz)], dtype=torch.float64, device="cuda:%d" % local)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        total_ms, nnz_all = float(tmax[0]), float(t[1])
    else:
        nnz_all = float(nnz)
    sampler.stop()

    # end to end through the C ABI with HOST buffers (pinned): every SpMV uploads its x and downloads
    # its y inside the timed region. (a) the reference's strictly synchronous sequence
    # upload -> spmv -> download+finish (sw/host.cpp:293-371); (b) the same three calls with the
    # asynchronous download, so that upload(k+1), SpMV(k) and download(k-1) overlap on the library's
    # copy streams (two host buffers in rotation) -- the way a throughput-oriented caller uses it.
    # Two different vectors alternate, so that a stale x or y anywhere in the pipeline shows up in the result.
    xw2 = np.ascontiguousarray(xw[::-1])
    ctx.upload_vector(xw2)
    ctx.spmv()
    y2 = ctx.download_result()
    want = (y, y2)
    sum_abs = None
    if impl != "fixed":                 # sum_i |a_i x_i| per row, the scale of fp32 rounding noise (plain numpy)
        starts = np.minimum(ip2[:-1].astype(np.int64), max(nnz - 1, 0))
        empty = np.diff(ip2.astype(np.int64)) == 0
        sum_abs = []
        for xv in (xw, xw2):
            t = np.abs(data.astype(np.float64)) * np.abs(xv.view(np.float32).astype(np.float64)[indices])
            sa = np.add.reduceat(t, starts) if nnz else np.zeros(r2)
            sa[empty] = 0.0
            sum_abs.append(sa)

    def same(got, k):
        # fixed point: sums of non-negative integers are order independent -> identical words; float: the
        # row updates are fp32 atomics whose order varies from launch to launch -> tolerance only
        if impl == "fixed":
            ok = np.array_equal(got, want[k])
        else:
            a, b = got.view(np.float32).astype(np.float64), want[k].view(np.float32).astype(np.float64)
            ok = bool(np.all(np.abs(a - b) <= 2e-5 * sum_abs[k] + 1e-30))   # each run is within 1e-5 * sum|a_i x_i|
        if not ok:
            ctx.close()
            raise SystemExit("bench: end-to-end result %d differs from the device-resident run" % k)

    px = [capi.PinnedArray(c2) for _ in range(2)]
    py = [capi.PinnedArray(r2) for _ in range(2)]
    px[0].array[:] = xw
    px[1].array[:] = xw2
    xh, yh = [b_.array for b_ in px], [b_.array for b_ in py]
    n_e2e = max(64, min(args.steps * B, 4096)) if WORKLOAD != "c5s" else 16
    n_e2e += n_e2e & 1
    ctx.time_e2e(xh, yh, 16, async_download=False)
    barrier()
    # (a) strictly synchronous, (b) pipelined; both issued from C by hsb_time_e2e through the public entry points
    n_sync = max(64, n_e2e // 4) if WORKLOAD != "c5s" else 8
    e2e_sync_s = ctx.time_e2e(xh, yh, n_sync + (n_sync & 1), async_download=False)
    same(yh[0], 0)
    same(yh[1], 1)
    for b_ in yh:
        b_[:] = 0
    barrier()
    e2e_s = ctx.time_e2e(xh, yh, n_e2e, async_download=True)
    same(yh[0], 0)
    same(yh[1], 1)
    # the same pipelined sequence issued call by call from Python (ctypes): what the pytest glue sees
    barrier()
    t0 = time.perf_counter()
    n_py = 256 if WORKLOAD != "c5s" else 8
    for k in range(n_py):
        ctx.upload_vector(xh[k & 1]); ctx.spmv(); ctx.download_result_async(yh[k & 1])
    ctx.sync()
    e2e_py_s = (time.perf_counter() - t0) / n_py
    same(yh[0], 0)
    same(yh[1], 1)
    if dist is not None:
        import torch
        t = torch.tensor([e2e_s, e2e_sync_s], dtype=torch.float64, device="cuda:%d" % local)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s, e2e_sync_s = float(t[0]), float(t[1])

    # the checker: the result words of this very run against the oracle's closed form -- on EVERY rank (each has
    # its own matrix or its own row block)
    from oracle import hsoracle
    port = hsoracle.Port()
    if impl == "fixed":
        ok = bool(np.array_equal(y, port.spmv_q824(ip2, indices, words, xw)))
        parity = "bit-exact vs oracle (closed form of ap_ufixed<32,8,AP_RND,AP_SAT> as written in oracle/shim/ap_fixed.h)"
    else:
        y64, sa = port.spmv_f64(ip2, indices, data, x)
        ok = bool(np.all(np.abs(y.view(np.float32).astype(np.float64) - y64) <= 1e-5 * sa + 1e-30))
        parity = "within 1e-5 * sum|a_i x_i| of the oracle's fp64 SpMV"
    parity += " checked in this run" + (" on every rank" if world > 1 else "")
    if dist is not None:
        ok = _max_over_ranks(dist, "cuda:%d" % local, 0.0 if ok else 1.0)[0] == 0.0
    if not ok:
        raise SystemExit("bench: GPU result differs from the oracle -- refusing to report a number")
    ctx.close()
    ctx = None
    sharded = None
    if world > 1 and WORKLOAD == "c2" and not SHARD_ONE_MATRIX:
        sharded = sharded_block(args, dist, rank, local, world)

    if rank == 0:
        peak, peak_src = load_peaks()
        sec_per_spmv = total_ms / 1e3 / (args.steps * B)
        gops = 2.0 * nnz_all / sec_per_spmv / 1e9
        alg = st["algorithmic_bytes"]
        achieved = alg / (kernel_ms / 1e3) / 1e9
        # DRAM bytes per launch from an ncu capture -- quoted only when that capture was taken on THIS build of the
        # library (profiles/traffic.json carries the SHA-256 of the profiled library's sources; tools/ncu_traffic.py writes it)
        traffic, traffic_note = None, "no ncu capture of this build (profiles/traffic.json was taken on other library sources)"
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp) and WORKLOAD == "c2" and impl == "fixed":
            tj = json.load(open(tp))
            if tj.get("source_sha256") == capi.source_hash():
                traffic = tj.get("dram_bytes_per_launch")
                traffic_note = {k: tj[k] for k in ("source", "range", "isolated_launch_us") if k in tj}
        out = {
            "metric": "SpMV GOPS (2*nnz/t, sw/benchmark.cpp:312-346)", "value": gops, "unit": "GOPS",
            "gbps": 8.0 * nnz_all / 2 ** 30 / sec_per_spmv,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "ms_per_spmv": 1e3 * sec_per_spmv, "higher_is_better": True,
            "scaling": "strong" if (SHARD_ONE_MATRIX and world > 1) else "weak", "vs_baseline": None,
            "dtype": "u32 Q8.24 (ap_ufixed<32,8,AP_RND,AP_SAT>), 64-bit accumulate" if impl == "fixed" else "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(nnz), "spmv_per_step": B, "l2_policy": "inputs larger than L2: %d HBM replicas of the matrix "
                       "(%.0f MB each) used round-robin" % (replicas, st["format_bytes"] / 1e6),
                       "sharding": (("nnz-balanced row-block shards of ONE matrix" if SHARD_ONE_MATRIX else
                                     "one matrix of the workload's size per GPU") +
                                    ", x replicated (one NCCL broadcast before timing), no data-path collective")
                       if world > 1 else "single GPU",
                       "tile_cols": st["tile_cols"], "col_tiles": st["n_col_tiles"], "grid": st["grid"]},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_src, "kernel": "spmv_tiles_kernel<%s>" % ("FixedArith" if impl == "fixed" else "FloatArith"),
                         "kernel_ms": kernel_ms, "kernel_ms_isolated": kernel_ms_isolated,
                         "algorithmic_bytes_per_launch": alg,
                         "format_bytes_per_launch": st["format_bytes"],
                         "format_frac": st["format_bytes"] / (kernel_ms / 1e3) / 1e9 / peak,
                         "note": "achieved = ALGORITHMIC bytes (8*nnz+4*(rows+1)+4*rows+4*cols) / mean launch duration "
                                 "in the timed loop. The kernel streams a 6 B/slot compressed format (16-bit tile-local "
                                 "columns), so frac can exceed 1.0 by compression; format_frac uses the bytes actually "
                                 "streamed"},
            "e2e": {"value": 2.0 * nnz_all / e2e_s / 1e9, "unit": "GOPS", "h2d_bytes_per_step": B * c2 * 4,
                    "d2h_bytes_per_step": B * r2 * 4, "ms_per_spmv": 1e3 * e2e_s,
                    "synchronous_value": 2.0 * nnz_all / e2e_sync_s / 1e9, "synchronous_ms_per_spmv": 1e3 * e2e_sync_s,
                    "python_glue_ms_per_spmv": 1e3 * e2e_py_s,
                    "what": "per SpMV: hsb_upload_vector(pinned x) + hsb_spmv + hsb_download_result_async(pinned y) "
                            "issued from C (hsb_time_e2e), two x / y host buffers alternating, copies and kernels "
                            "overlapping across consecutive SpMVs, wall clock incl. the final hsb_sync; synchronous_* = "
                            "the same with the blocking hsb_download_result after every SpMV; python_glue_* = the "
                            "pipelined sequence issued call by call through ctypes; matrix resident as in "
                            "sw/benchmark.cpp"},
            "gpu_launches": int(launches2 - launches1),
            "gpu_launches_per_spmv": (launches2 - launches1) / float(args.steps * B),
            "clocks": sampler.summary(),
            "preprocess_s": st["preprocess_seconds"],
        }
        out["config"]["parity"] = parity
        if placement is not None:
            out["config"]["host_placement"] = {"rank0_gpu_numa_node": placement[0], "rank0_cpus_pinned": placement[1],
                                               "what": "every rank's host thread runs on the CPUs of its GPU's NUMA node (sysfs); "
                                                       "page-locked buffers are allocated on that node (hsb_host_alloc)"}
        if sharded is not None:
            out["sharded"] = sharded
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(r2, c2, ip2, indices, data, x, nnz)
        emit(out)
    if ctx is not None:
        ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
