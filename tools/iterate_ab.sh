#!/bin/bash
# same-box A/B of hsb_iterate: one cooperative launch against a launch per step (tests/pagerank.py)
cd "$(dirname "$0")/.."
for cfg in "--nodes 5120 --nnz 170000" "--nodes 107614 --nnz 13670000" "--nodes 576289 --nnz 42460000"; do
  for impl in fixed float_pob; do
    for form in "" "--step-form"; do
      echo -n "$cfg $impl $form: "
      timeout 300 python tests/pagerank.py $cfg --impl $impl --iters 200 --check $form 2>&1 | tail -1 | python -c "import sys,json; d=json.load(sys.stdin); print(round(d['ms_per_iteration']*1e3,2), 'us', d.get('parity'))"
    done
  done
done
