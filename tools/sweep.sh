export PYTHONUNBUFFERED=1
cd "$(dirname "$0")/.."
for w in c2 c4 c3 c1; do
  for v in nohint ""; do
    if [ -z "$v" ]; then SWEEP_WORKLOAD=$w timeout 200 python tools/sweep_time.py 2>&1 | tail -1
    else SWEEP_WORKLOAD=$w HSB_LIB=$PWD/hisparse_b200/libhsb_$v.so timeout 200 python tools/sweep_time.py 2>&1 | tail -1; fi
  done
done
for impl in float_pob fixed; do
  HSB_LIB=$PWD/hisparse_b200/libhsb_nohint.so timeout 200 python tools/c5_probe.py --impl $impl --no-check 2>&1 | tail -1 | cut -c1-250
  timeout 200 python tools/c5_probe.py --impl $impl 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('hints', d['impl'], d['ms_per_spmv'], d['gops'], d.get('parity'))"
done
