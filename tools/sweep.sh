#!/bin/bash
# A/B of library builds and run-time knobs on the resident-matrix loop.
#   tools/sweep.sh "c2 c4" "default nohint pf5"      (tags: hisparse_b200/libhsb_<tag>.so from tools/build_variants.sh)
#   HSB_SLICE_COST=6 tools/sweep.sh c2               (environment knobs are passed through)
export PYTHONUNBUFFERED=1
cd "$(dirname "$0")/.."
for w in ${1:-c2}; do
  for v in ${2:-default}; do
    if [ "$v" = default ]; then SWEEP_WORKLOAD=$w timeout 300 python tools/sweep_time.py 2>&1 | tail -1
    else SWEEP_WORKLOAD=$w HSB_LIB=$PWD/hisparse_b200/libhsb_$v.so timeout 300 python tools/sweep_time.py 2>&1 | tail -1; fi
  done
done
