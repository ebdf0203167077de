# 4-GPU pass: weak-scaling bench line, C4 row-block sharded (strong scaling), multi-GPU iteration over peer memory
export PYTHONUNBUFFERED=1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-4}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"
timeout 600 $T bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/b$N.err | tail -1 > gpurun_out/bench_n$N.json
timeout 600 $T bench.py --gpus $N --steps 10 --warmup 3 --workload c4 --shard-one-matrix 2>gpurun_out/b${N}c4.err | tail -1 > gpurun_out/bench_n${N}_c4_sharded.json
for f in bench_n$N bench_n${N}_c4_sharded; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/$f.json"))
    print("$f", "gpus", d["n_gpus"], "us/spmv %.2f" % (1e3*d["ms_per_spmv"]), "GOPS %.0f" % d["value"], "e2e %.0f" % d["e2e"]["value"], d["scaling"], "frac %.3f" % d["roofline"]["frac"])
except Exception as e:
    print("$f", "FAILED", e, open("gpurun_out/$f.json").read()[:300]); print(open("gpurun_out/" + ("b$N.err" if "$f" == "bench_n$N" else "b${N}c4.err")).read()[-800:])
PY
done
timeout 300 $T tests/pagerank.py --iters 50 --p2p 2>&1 | tail -1
timeout 300 $T tests/pagerank.py --iters 50 --impl fixed --nodes 107614 --nnz 13670000 --check --p2p 2>&1 | tail -1
