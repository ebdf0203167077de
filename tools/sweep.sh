export PYTHONUNBUFFERED=1
cd "$(dirname "$0")/.."
for w in c3 c2 c4 c1; do
  for v in norot "" bp4k bp8k bp32k; do
    if [ -z "$v" ]; then SWEEP_WORKLOAD=$w timeout 200 python tools/sweep_time.py 2>&1 | tail -1
    else SWEEP_WORKLOAD=$w HSB_LIB=$PWD/hisparse_b200/libhsb_$v.so timeout 200 python tools/sweep_time.py 2>&1 | tail -1; fi
  done
done
