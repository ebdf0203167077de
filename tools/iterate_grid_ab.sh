#!/bin/bash
# grid size of a small matrix (HSB_MIN_STEPS_PER_CTA) and the one-launch iteration: fewer CTAs = cheaper grid barriers?
cd "$(dirname "$0")/.."
for cfg in "--nodes 5120 --nnz 170000" "--nodes 26904 --nnz 3417500"; do
  for m in "" 16 64 256 1024; do
    if [ -z "$m" ]; then unset HSB_MIN_STEPS_PER_CTA; else export HSB_MIN_STEPS_PER_CTA=$m; fi
    echo -n "$cfg fixed [min steps per CTA ${m:-default}]: "
    timeout 300 python tests/pagerank.py $cfg --impl fixed --iters 500 2>&1 | tail -1 | python -c "import sys,json; d=json.load(sys.stdin); print(round(d['ms_per_iteration']*1e3,2), 'us')"
  done
done
