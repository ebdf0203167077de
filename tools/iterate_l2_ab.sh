#!/bin/bash
# does the iteration gain from keeping the matrix in the L2 (no evict-first hint on the stream)?  same box, C2-sized and a 60 MB matrix
cd "$(dirname "$0")/.."
for cfg in "--nodes 107614 --nnz 13670000" "--nodes 80000 --nnz 8000000"; do
  for lib in "" nohints "" nohints; do
    if [ -z "$lib" ]; then unset HSB_LIB; else export HSB_LIB=$PWD/hisparse_b200/libhsb_$lib.so; fi
    echo -n "$cfg fixed [${lib:-shipped}]: "
    timeout 300 python tests/pagerank.py $cfg --impl fixed --iters 300 2>&1 | tail -1 | python -c "import sys,json; d=json.load(sys.stdin); print(round(d['ms_per_iteration']*1e3,2), 'us')"
  done
done
