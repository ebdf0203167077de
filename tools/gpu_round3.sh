# 2-GPU pass: multi-GPU iteration, NCCL-from-host versus peer-memory fused exchange
export PYTHONUNBUFFERED=1
cd "$(dirname "$0")/.."
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 300 $T tests/pagerank.py --iters 50 --impl fixed --nodes 107614 --nnz 13670000 --check 2>&1 | tail -1
timeout 300 $T tests/pagerank.py --iters 50 --impl fixed --nodes 107614 --nnz 13670000 --check --p2p 2>&1 | tail -2
timeout 300 $T tests/pagerank.py --iters 50 --p2p 2>&1 | tail -1
timeout 300 $T tests/pagerank.py --iters 50 2>&1 | tail -1
timeout 200 python tests/pagerank.py --iters 50 2>&1 | tail -1
