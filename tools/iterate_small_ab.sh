#!/bin/bash
# small per-GPU work: world of N ranks, matrix scaled so that each rank holds ~3.4 M non-zeros (a quarter of C2)
cd "$(dirname "$0")/.."
N=${1:-1}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
nodes=$((26904 * N)); nnz=$((3417500 * N))
for form in "" "--step-form" "" "--step-form"; do
  echo -n "$N GPUs nodes $nodes nnz $nnz fixed $form: "
  timeout 300 $T tests/pagerank.py --nodes $nodes --nnz $nnz --impl fixed --iters 300 --p2p --check $form 2>&1 | grep '^{' | tail -1 | python -c "import sys,json; d=json.load(sys.stdin); print(round(d['ms_per_iteration']*1e3,2), 'us', d.get('parity'))"
done
