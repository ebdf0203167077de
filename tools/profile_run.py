"""Small driver for ncu: uploads the bench workload (or a named config) and runs a few SpMVs.
    ncu ... python tools/profile_run.py --spmv 8 [--impl fixed|float_pob] [--config c2|c1|c3|c4]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hisparse_b200 import capi, matgen  # noqa: E402


def make(config):
    if config == "c1":
        return matgen.random_csr(4096, 4096, 0.01, 0xC0FFEE01)
    if config == "c3":
        return matgen.bernoulli_csr(512, 33288, 0.5, 0xC0FFEE03)
    if config == "t95":
        return matgen.bernoulli_csr(512, 33288, 0.05, 0xC0FFEE03)
    if config == "c4":
        return matgen.rmat_csr(576289, 42_460_000, 0xC0FFEE04, symmetric=True, oversample=1.5)
    return matgen.rmat_csr(107614, 13_670_000, 0xC0FFEE02)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spmv", type=int, default=8)
    ap.add_argument("--impl", default="fixed")
    ap.add_argument("--config", default="c2")
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--replicas", type=int, default=0)
    ap.add_argument("--gpu-format", action="store_true")
    ap.add_argument("--range", type=int, default=0,
                    help="N back-to-back SpMVs inside ONE cudaProfilerStart/Stop range (ncu --replay-mode app-range)")
    a = ap.parse_args()
    rows, cols, indptr, indices, data = make(a.config)
    r2, c2, ip2 = matgen.pad_csr(rows, cols, indptr, 128, 8)
    x = np.zeros(c2, np.float32)
    x[:cols] = np.random.default_rng(1).random(cols, dtype=np.float32)
    if a.impl == "fixed":
        data, x = matgen.quantize_q824(data * np.float32(0.05)), matgen.quantize_q824(x)
    ctx = capi.Context(0, a.impl)
    import time as _t
    t0 = _t.perf_counter()
    ctx.upload_matrix_csr(r2, c2, ip2, indices, data, on_gpu=a.gpu_format)
    t_up = _t.perf_counter() - t0
    st = ctx.stats()
    ctx.set_replicas(a.replicas or max(2, int(np.ceil(2.5 * (capi.device_l2_bytes(0) or 126 * 2 ** 20) / max(st["format_bytes"], 1)))))
    ctx.upload_vector(x)
    if a.time:
        step, kern = ctx.time_spmv(20, 400)
        print("config %s impl %s nnz %d: %.2f us/spmv, kernel %.2f us, %.0f GOPS, alg %.0f GB/s, fmt %.0f GB/s" % (
            a.config, a.impl, st["nnz"], step * 1e3, kern * 1e3, 2 * st["nnz"] / step / 1e6,
            st["algorithmic_bytes"] / kern / 1e6, st["format_bytes"] / kern / 1e6))
        print("upload+format wall %.3f s, preprocess %.4f s (%s)" % (t_up, st["preprocess_seconds"], "GPU" if a.gpu_format else "host"))
        print(st)
    elif a.range:
        for _ in range(64):
            ctx.spmv()
        ctx.profiler(True)
        for _ in range(a.range):
            ctx.spmv()
        ctx.profiler(False)
        print("range: %d SpMVs, algorithmic bytes %d, format bytes %d per SpMV" % (a.range, st["algorithmic_bytes"], st["format_bytes"]))
    else:
        for _ in range(a.spmv):
            ctx.spmv()
        ctx.sync()
    ctx.close()


if __name__ == "__main__":
    main()
