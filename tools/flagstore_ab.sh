#!/bin/bash
# same-box A/B: arrival flags published by fence + relaxed stores (shipped) against st.release.sys per target (libhsb_relst.so)
cd "$(dirname "$0")/.."
N=${1:-2}
for rep in 1 2; do
  for lib in "" relst; do
    if [ -z "$lib" ]; then unset HSB_LIB; else export HSB_LIB=$PWD/hisparse_b200/libhsb_$lib.so; fi
    echo -n "c2 resident [${lib:-shipped}]: "; python tools/abi_time.py c2 2>&1 | tail -1
  done
done
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
for lib in "" relst; do
  if [ -z "$lib" ]; then unset HSB_LIB; else export HSB_LIB=$PWD/hisparse_b200/libhsb_$lib.so; fi
  for form in "" "--step-form"; do
    echo -n "$N GPUs C2-sized fixed [${lib:-shipped}] $form: "
    timeout 400 $T tests/pagerank.py --nodes 107614 --nnz 13670000 --impl fixed --iters 200 --p2p --check $form 2>&1 | grep '^{' | tail -1 | python -c "import sys,json; d=json.load(sys.stdin); print(round(d['ms_per_iteration']*1e3,2), 'us', d.get('parity'))"
  done
done
