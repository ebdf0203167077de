# one GPU-box pass (1 GPU): parity tests, smoke, the bench line, C4
export PYTHONUNBUFFERED=1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --no-cpu-baseline 2>gpurun_out/bench_c2.err | tail -1 > gpurun_out/bench_c2b.json; tail -c 300 gpurun_out/bench_c2.err
timeout 300 python bench.py --workload c4 --no-cpu-baseline --steps 10 2>&1 | tail -1 > gpurun_out/bench_c4b.json
for w in c2b c4b; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$w.json"))
    print("$w", "us/spmv %.2f" % (1e3*d["ms_per_spmv"]), "GOPS %.0f" % d["value"], "e2e %.0f (%.2f us)" % (d["e2e"]["value"], 1e3*d["e2e"]["ms_per_spmv"]), "frac %.3f fmt %.3f" % (d["roofline"]["frac"], d["roofline"]["format_frac"]), d["config"]["col_tiles"], d["config"]["tile_cols"], d["clocks"]["reasons"])
except Exception as e:
    print("$w FAILED", e, open("gpurun_out/bench_$w.json").read()[:300])
PY
done
