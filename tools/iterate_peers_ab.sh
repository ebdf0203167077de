#!/bin/bash
# same-box A/B of the multi-GPU iteration: one resident kernel per GPU against a launch per step (tests/pagerank.py --p2p)
# usage: tools/iterate_peers_ab.sh <n_gpus>
cd "$(dirname "$0")/.."
N=${1:-2}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
for cfg in "--nodes 107614 --nnz 13670000" "--nodes 576289 --nnz 42460000"; do
  for impl in ${IMPLS:-fixed float_pob}; do
    for form in "" "--step-form"; do
      echo -n "$N GPUs $cfg $impl $form: "
      timeout 400 $T tests/pagerank.py $cfg --impl $impl --iters 200 --p2p --check $form 2>&1 | grep '^{' | tail -1 | python -c "import sys,json; d=json.load(sys.stdin); print(round(d['ms_per_iteration']*1e3,2), 'us', d.get('parity'))"
    done
  done
done
