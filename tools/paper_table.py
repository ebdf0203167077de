"""Table 3 of the HiSparse paper (throughput in GOPS and bandwidth efficiency in MOPS per GB/s of peak memory
bandwidth, fixed-point design) re-measured with this engine on one B200, next to the published MKL / cuSPARSE /
HiSparse(U280) columns, plus the preprocessing times of Table 8.

    python tools/paper_table.py [--datasets DIR] [--out gpurun_out/paper_table.md]

The datasets are downloads the reference does not ship (datasets/download.sh). If DIR holds the reference's
`.npz` files (scipy CSR: indptr / indices / data / shape) they are used; otherwise stand-ins of the published
size and density are generated: Bernoulli masks for the pruned transformer layers, R-MAT for the graphs up to
ogbl-ppa, and the device-side power-law generator (uniform columns) for the three largest graphs. Like the
reference's benchmark (sw/benchmark.cpp:57-59) values are irrelevant to the timing; the fixed-point path is used."""
import argparse
import json
import os
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# name, file stem (sw/bm.sh), rows, cols, density, published GOPS (MKL, cuSPARSE, HiSparse), published preprocessing s
DATASETS = [
    ("transformer-50", "transformer_50_512_33288", 512, 33288, 0.50, (5.9, 26.9, 21.9), 0.24),
    ("transformer-60", "transformer_60_512_33288", 512, 33288, 0.40, (5.6, 21.5, 18.9), 0.16),
    ("transformer-70", "transformer_70_512_33288", 512, 33288, 0.30, (5.2, 17.7, 16.5), 0.11),
    ("transformer-80", "transformer_80_512_33288", 512, 33288, 0.20, (4.1, 19.4, 14.8), 0.08),
    ("transformer-90", "transformer_90_512_33288", 512, 33288, 0.10, (2.3, 13.6, 9.7), 0.04),
    ("transformer-95", "transformer_95_512_33288", 512, 33288, 0.05, (1.2, 10.7, 5.7), 0.02),
    ("mouse-gene", "mouse_gene_45K_29M", 45101, 45101, 1.42e-2, (12.1, 29.0, 27.2), 0.87),
    ("googleplus", "gplus_108K_13M", 107614, 107614, 1.2e-3, (5.1, 27.2, 21.2), 0.39),
    ("ogbl-ppa", "ogbl_ppa_576K_42M", 576289, 576289, 127.9e-6, (4.1, 18.0, 24.4), 1.89),
    ("hollywood", "hollywood_1M_113M", 1069126, 1069126, 98.5e-6, (4.4, 22.6, 24.9), 4.68),
    ("pokec", "pokec_1633K_31M", 1632803, 1632803, 11.5e-6, (3.0, 10.5, 11.2), 3.43),
    ("ogbn-products", "ogbn_products_2M_124M", 2449029, 2449029, 20.6e-6, (3.1, 5.0, 20.6), 10.60),
]
PEAK_GBS = {"MKL (2x Xeon 6242)": 282.0, "cuSPARSE (GTX 1080 Ti)": 484.0, "HiSparse (U280)": 258.0, "B200 (HBM3e spec)": 8000.0}


def find_npz(directory, stem):
    if not directory:
        return None
    for root, _, files in os.walk(directory):
        for f in files:
            if f.startswith(stem) and f.endswith(".npz"):
                return os.path.join(root, f)
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--datasets", default=os.path.join(ROOT, "datasets"))
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "paper_table.md"))
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    from hisparse_b200 import capi, matgen
    rows_out = []
    for name, stem, rows, cols, density, pub, pub_prep in DATASETS:
        if args.only and args.only not in name:
            continue
        nnz_target = int(round(rows * cols * density))
        path = find_npz(args.datasets, stem)
        ctx = capi.Context(0, "fixed")
        t0 = time.perf_counter()
        if path:
            d = np.load(path)
            r, c = [int(v) for v in d["shape"]]
            ip, ix = d["indptr"].astype(np.uint32), d["indices"].astype(np.uint32)
            r2, c2, ip2 = matgen.pad_csr(r, c, ip, 128, 8)
            words = np.full(ix.size, 1 << 10, np.uint32)           # the reference overwrites the values too
            source = os.path.basename(path)
            t_gen = time.perf_counter() - t0
            t0 = time.perf_counter()
            ctx.upload_matrix_csr(r2, c2, ip2, ix, words)
            nnz = int(ix.size)
        elif rows * cols * density > 60e6 or rows > 1_000_000:
            r2, c2 = rows + (-rows) % 128, cols + (-cols) % 8
            m = capi.DeviceCsr.powerlaw(0, r2, c2, mean_degree=nnz_target / rows, alpha=2.1, max_degree=200000,
                                        band_half_width=1, band_fraction=0.0, seed=zlib.crc32(name.encode()), q824=True,
                                        value_scale=0.01)
            source = "power-law stand-in (device generator)"
            t_gen = time.perf_counter() - t0
            t0 = time.perf_counter()
            ctx.upload_matrix_csr_device(m)
            nnz = int(m.nnz)
            m.free()
        else:
            if rows == 512:
                r, c, ip, ix, data = matgen.bernoulli_csr(rows, cols, density, 0xC0FFEE03, values="small")
                source = "Bernoulli mask stand-in"
            elif density > 5e-3:
                r, c, ip, ix, data = matgen.random_csr(rows, cols, density, 0xC0FFEE06, values="small")
                source = "uniform random stand-in"
            else:
                r, c, ip, ix, data = matgen.rmat_csr(rows, nnz_target, 0xC0FFEE02 if name == "googleplus" else 0xC0FFEE04,
                                                     values="small", symmetric=name == "ogbl-ppa",
                                                     oversample=1.5 if name == "ogbl-ppa" else 1.36)
                source = "R-MAT stand-in"
            r2, c2, ip2 = matgen.pad_csr(r, c, ip, 128, 8)
            words = matgen.quantize_q824(data)
            t_gen = time.perf_counter() - t0
            t0 = time.perf_counter()
            ctx.upload_matrix_csr(r2, c2, ip2, ix, words)
            nnz = int(ix.size)
        t_up = time.perf_counter() - t0
        st = ctx.stats()
        ctx.set_replicas(max(2, int(np.ceil(2.5 * (capi.device_l2_bytes(0) or 126 * 2 ** 20) / max(st["format_bytes"], 1)))))
        ctx.upload_vector(np.full(c2, 1 << 24, np.uint32))
        n = max(50, min(2000, int(0.05 / max(1e-6, nnz * 8 / 5e12))))
        ms, _ = ctx.time_spmv(n // 5, n, kernel=False)
        ctx.close()
        gops = 2.0 * nnz / ms / 1e6
        rows_out.append({"name": name, "rows": r2, "cols": c2, "nnz": nnz, "source": source, "us_per_spmv": 1e3 * ms,
                         "gops": gops, "gbps": 8.0 * nnz / 2 ** 30 / (ms / 1e3), "bw_eff_b200": 1e3 * gops / PEAK_GBS["B200 (HBM3e spec)"],
                         "alg_frac_of_8TBs": st["algorithmic_bytes"] / (ms / 1e3) / 8e12,
                         "preprocess_s": st["preprocess_seconds"], "upload_wall_s": t_up, "generate_s": t_gen,
                         "published": pub, "published_preprocess_s": pub_prep})
        print(json.dumps(rows_out[-1]), flush=True)
    geo = lambda v: float(np.exp(np.mean(np.log(v)))) if len(v) else float("nan")
    lines = ["| dataset | rows x cols | nnz | source | MKL | cuSPARSE | HiSparse U280 | **B200 (this)** | us / SpMV | BW-eff HiSparse | **BW-eff B200** | % of 8 TB/s (alg. bytes) | preprocessing U280-host s | **here s** |",
             "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for r in rows_out:
        p = r["published"]
        lines.append("| %s | %d x %d | %.1f M | %s | %.1f | %.1f | %.1f | **%.0f** | %.1f | %.1f | **%.1f** | %.0f %% | %.2f | **%.3f** |" % (
            r["name"], r["rows"], r["cols"], r["nnz"] / 1e6, r["source"], p[0], p[1], p[2], r["gops"], r["us_per_spmv"],
            1e3 * p[2] / PEAK_GBS["HiSparse (U280)"], r["bw_eff_b200"], 100 * r["alg_frac_of_8TBs"], r["published_preprocess_s"],
            r["preprocess_s"]))
    if rows_out:
        lines.append("| geomean | | | | %.1f | %.1f | %.1f | **%.0f** | | %.1f | **%.1f** | | | |" % (
            geo([r["published"][0] for r in rows_out]), geo([r["published"][1] for r in rows_out]),
            geo([r["published"][2] for r in rows_out]), geo([r["gops"] for r in rows_out]),
            geo([1e3 * r["published"][2] / 258.0 for r in rows_out]), geo([r["bw_eff_b200"] for r in rows_out])))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    open(args.out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
