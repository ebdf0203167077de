# 2-GPU pass: weak-scaling bench line, C4 row-block sharded (strong scaling), multi-GPU PageRank
export PYTHONUNBUFFERED=1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $T bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/b2.err | tail -1 > gpurun_out/bench_n2.json; tail -c 400 gpurun_out/b2.err
timeout 600 $T bench.py --gpus 2 --steps 10 --warmup 3 --workload c4 --shard-one-matrix 2>gpurun_out/b2c4.err | tail -1 > gpurun_out/bench_n2_c4_sharded.json; tail -c 400 gpurun_out/b2c4.err
timeout 300 python bench.py --workload c3 --no-cpu-baseline --steps 10 2>&1 | tail -1 > gpurun_out/bench_c3.json
for f in bench_n2 bench_n2_c4_sharded bench_c3; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/$f.json"))
    print("$f", "gpus", d["n_gpus"], "us/spmv %.2f" % (1e3*d["ms_per_spmv"]), "GOPS %.0f" % d["value"], "e2e %.0f" % d["e2e"]["value"], d["scaling"], "frac %.3f" % d["roofline"]["frac"])
except Exception as e:
    print("$f", "FAILED", e, open("gpurun_out/$f.json").read()[:300])
PY
done
timeout 300 $T tests/pagerank.py --iters 20 2>&1 | tail -1
timeout 300 $T tests/pagerank.py --iters 20 --impl fixed --nodes 107614 --nnz 13670000 2>&1 | tail -1
