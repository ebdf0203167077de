"""Probe of BASELINE config C5 on ONE GPU: a row-block shard of the 100 M x 100 M power-law matrix is
generated on the device, formatted on the device, multiplied, and checked against the oracle on the
host (full size: the oracle's C loop does 250 M MACs in about a second).

    python tests/c5_probe.py [--rows 12500000] [--cols 100000000] [--impl fixed|float_pob] [--no-check]

Prints one JSON line per run. Not a bench line (bench.py is); used to fill DESIGN.md's C5 table."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=12_500_000)
    ap.add_argument("--cols", type=int, default=100_000_000)
    ap.add_argument("--first-row", type=int, default=0)
    ap.add_argument("--impl", default="fixed")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--rows-per-partition", type=int, default=0)
    ap.add_argument("--no-check", action="store_true")
    args = ap.parse_args()
    from hisparse_b200 import capi
    fixed = args.impl == "fixed"
    t0 = time.perf_counter()
    m = capi.DeviceCsr.powerlaw(0, args.rows, args.cols, first_global_row=args.first_row, q824=fixed,
                                value_scale=0.05 if fixed else 1.0)
    t_gen = time.perf_counter() - t0
    ctx = capi.Context(0, args.impl)
    t0 = time.perf_counter()
    ctx.upload_matrix_csr_device(m, args.rows_per_partition)
    t_fmt = time.perf_counter() - t0
    st = ctx.stats()
    rng = np.random.default_rng(5)
    xf = rng.random(args.cols, dtype=np.float32)
    if fixed:
        from hisparse_b200 import matgen
        xw = matgen.quantize_q824(xf)
    else:
        xw = xf.view(np.uint32)
    ctx.upload_vector(xw)
    ctx.spmv()
    y = ctx.download_result()
    step_ms, _ = ctx.time_spmv(2, args.steps, kernel=False)
    _, isolated_ms = ctx.time_spmv(1, args.steps, kernel=True)      # an event pair around every launch: no overlap between launches
    out = {"rows": args.rows, "cols": args.cols, "nnz": int(m.nnz), "impl": args.impl, "generate_s": t_gen,
           "format_s": t_fmt, "ms_per_spmv": step_ms, "ms_isolated_launch": isolated_ms, "gops": 2.0 * m.nnz / step_ms / 1e6,
           "alg_gbs": st["algorithmic_bytes"] / step_ms / 1e6, "format_gbs": st["format_bytes"] / step_ms / 1e6,
           "format_bytes_per_nnz": st["format_bytes"] / max(m.nnz, 1), "n_col_tiles": st["n_col_tiles"],
           "tile_cols": st["tile_cols"], "n_streams": st["n_streams"], "n_slices": st["n_slices"],
           "n_elems": st["n_elems"], "layout": st.get("layout")}
    if not args.no_check:
        from oracle import hsoracle
        port = hsoracle.Port()
        ip, ix, vv = m.download()
        assert ip[0] == 0 and ip[-1] == m.nnz and np.all(np.diff(ip.astype(np.int64)) >= 0)
        t0 = time.perf_counter()
        if fixed:
            want = port.spmv_q824(ip, ix, vv, xw)
            out["parity"] = "bit-exact" if np.array_equal(y, want) else "MISMATCH in %d rows" % int((y != want).sum())
        else:
            y64, sa = port.spmv_f64(ip, ix, vv.view(np.float32), xf)
            ok = np.all(np.abs(y.view(np.float32).astype(np.float64) - y64) <= 1e-5 * sa + 1e-30)
            out["parity"] = "within 1e-5 * sum|a x|" if ok else "OUT OF TOLERANCE"
        out["oracle_s"] = time.perf_counter() - t0
        deg = np.diff(ip.astype(np.int64))
        out["degree"] = {"mean": float(deg.mean()), "max": int(deg.max()), "zero_rows": int((deg == 0).sum())}
    print(json.dumps(out), flush=True)
    ctx.close()
    m.free()


if __name__ == "__main__":
    main()
