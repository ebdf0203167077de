#!/bin/bash
# x-tile staging piece size (bytes per cp.async.bulk) and the iteration, where every CTA stages at the same moment
cd "$(dirname "$0")/.."
for lib in "" bp4 bp8 bp32 bp64 ""; do
  if [ -z "$lib" ]; then unset HSB_LIB; else export HSB_LIB=$PWD/hisparse_b200/libhsb_$lib.so; fi
  echo -n "C2-sized fixed [${lib:-shipped 16 KB}]: "
  timeout 300 python tests/pagerank.py --nodes 107614 --nnz 13670000 --impl fixed --iters 300 2>&1 | tail -1 | python -c "import sys,json; d=json.load(sys.stdin); print(round(d['ms_per_iteration']*1e3,2), 'us')"
done
