# one GPU-box pass (1 GPU): parity tests, smoke, the bench line (+ reference arm), ncu launch list + full capture
export PYTHONUNBUFFERED=1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 | tail -1 > gpurun_out/bench_ref.json
timeout 600 python bench.py 2>gpurun_out/bench_c2.err | tail -1 > gpurun_out/bench_c2.json; tail -c 300 gpurun_out/bench_c2.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01f_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --batch 64 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmv_tiles --launch-skip 20 -c 1 -f -o gpurun_out/r01f_c2_fixed \
    python tools/profile_run.py --spmv 30 > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
for w in c1 c3 c4; do timeout 300 python bench.py --workload $w --no-cpu-baseline --steps 10 2>&1 | tail -1 > gpurun_out/bench_$w.json; done
for w in ref c2 c1 c3 c4; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$w.json"))
    if "$w" == "ref":
        print("ref", d["value"], d["cpu_baseline"])
    else:
        print("$w", "us/spmv %.2f" % (1e3*d["ms_per_spmv"]), "GOPS %.0f" % d["value"], "e2e %.0f (%.2f us)" % (d["e2e"]["value"], 1e3*d["e2e"]["ms_per_spmv"]), "frac %.3f fmt %.3f" % (d["roofline"]["frac"], d["roofline"]["format_frac"]), "cpu", d.get("cpu_baseline",{}).get("value"), d["clocks"]["reasons"])
except Exception as e:
    print("$w FAILED", e, open("gpurun_out/bench_$w.json").read()[:300])
PY
done
