"""Turn ncu CSV output into profiles/traffic.json, keyed by the SHA-256 of the library sources that were profiled (capi.source_hash), so
that bench.py only quotes DRAM traffic measured on the build it is running.
    python tools/ncu_traffic.py --kernel-csv K.csv [--range-csv R.csv --range-spmvs 512] --out profiles/traffic.json
K.csv: `ncu --csv --page raw` of one launch of spmv_tiles_kernel (--set full or the dram metrics);
R.csv: `ncu --csv --page raw --replay-mode app-range` of tools/profile_run.py --range N."""
import argparse
import csv
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rows_of(path):
    text = open(path, errors="replace").read()
    start = text.find('"ID"')
    if start < 0:
        return []
    rd = list(csv.reader(io.StringIO(text[start:])))
    head, units = rd[0], rd[1]
    out = []
    for r in rd[2:]:
        if len(r) == len(head):
            out.append({h: (v, u) for h, v, u in zip(head, r, units)})
    return out


def num(cell, want_unit_bytes=False):
    v, u = cell
    f = float(v.replace(",", ""))
    if want_unit_bytes:
        f *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
    return f


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kernel-csv")
    ap.add_argument("--range-csv")
    ap.add_argument("--range-spmvs", type=int, default=512)
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "traffic.json"))
    a = ap.parse_args()
    sys.path.insert(0, ROOT)
    from hisparse_b200 import capi
    out = {"workload": a.workload, "source_sha256": capi.source_hash()}
    if a.kernel_csv:
        ks = [r for r in rows_of(a.kernel_csv) if "spmv_tiles_kernel" in r.get("Kernel Name", ("", ""))[0]]
        if ks:
            k = ks[-1]
            rd, wr = num(k["dram__bytes_read.sum"], True), num(k["dram__bytes_write.sum"], True)
            out.update({"kernel": k["Kernel Name"][0], "dram_read": rd, "dram_write": wr, "dram_bytes_per_launch": rd + wr,
                        "isolated_launch_us": num(k["gpu__time_duration.sum"]) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(k["gpu__time_duration.sum"][1], 1),
                        "source": "ncu kernel replay of one launch (serialised, cold L2): " + os.path.basename(a.kernel_csv)})
    if a.range_csv:
        rs = rows_of(a.range_csv)
        if rs:
            r = rs[-1]
            rd, wr = num(r["dram__bytes_read.sum"], True), num(r["dram__bytes_write.sum"], True)
            t_us = num(r["gpu__time_duration.sum"]) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(r["gpu__time_duration.sum"][1], 1)
            out["range"] = {"spmvs": a.range_spmvs, "dram_read": rd, "dram_write": wr, "duration_us": t_us,
                            "dram_bytes_per_spmv": (rd + wr) / a.range_spmvs, "us_per_spmv": t_us / a.range_spmvs,
                            "dram_bytes_per_second": (rd + wr) / (t_us * 1e-6),
                            "source": "ncu --replay-mode app-range over %d back-to-back, overlapping launches: %s"
                                      % (a.range_spmvs, os.path.basename(a.range_csv))}
    json.dump(out, open(a.out, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
