#!/bin/bash
# fence.acq_rel (shipped) against fence.sc (libhsb_scf.so) in the barriers of the resident iteration kernel
# usage: tools/iterate_fence_ab.sh <n_gpus>
cd "$(dirname "$0")/.."
N=${1:-1}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
for cfg in "--nodes 5120 --nnz 170000" "--nodes 107614 --nnz 13670000"; do
  for lib in "" scf "" scf; do
    if [ -z "$lib" ]; then unset HSB_LIB; else export HSB_LIB=$PWD/hisparse_b200/libhsb_$lib.so; fi
    echo -n "$N GPUs $cfg fixed [${lib:-shipped acq_rel}]: "
    if [ "$N" = 1 ]; then timeout 300 python tests/pagerank.py $cfg --impl fixed --iters 400 --check 2>&1 | tail -1
    else timeout 300 $T tests/pagerank.py $cfg --impl fixed --iters 400 --p2p --check 2>&1 | grep '^{' | tail -1; fi | python -c "import sys,json; d=json.load(sys.stdin); print(round(d['ms_per_iteration']*1e3,2), 'us', d.get('parity'))"
  done
done
