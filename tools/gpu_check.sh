# one GPU-box pass: parity tests, smoke, bench (flag pipeline on / off)
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 400 python bench.py --steps 10 2>&1 | tail -1 | tee gpurun_out/bench_c2.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('FLAGS  value', d['value'], 'ms', d['ms_per_spmv'], 'e2e', d['e2e'], 'clocks', d['clocks'], 'cpu', d.get('cpu_baseline'))"
HSB_NO_FLAGS=1 timeout 400 python bench.py --steps 10 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_c2_noflags.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('EVENTS value', d['value'], 'ms', d['ms_per_spmv'], 'e2e', d['e2e'])"
