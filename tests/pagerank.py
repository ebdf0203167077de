"""PageRank-style power iteration x <- alpha (*) A x (+) beta on 1..N GPUs (the
caller either side of the SpMV in the reference's intended use, unit_tests/test_app.cpp:51-136).

    python tests/pagerank.py [--nodes 576289 --nnz 42460000 --iters 20 --impl float_pob]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/pagerank.py ...

One process per GPU. The link matrix (R-MAT, value 1/out-degree) is cut into nnz-balanced row blocks;
every rank keeps its block resident plus a replica of x. Per iteration: hsb_spmv on the block,
hsb_axpb_to_vector writes alpha*y+beta into the next x buffer at the rank's row offset (fused with the row
drain), the blocks are all-gathered IN PLACE over NCCL (device pointers from hsb_device_x_next), and
hsb_vector_commit flips the buffers. Nothing visits the host. Prints one JSON line (rank 0)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class _Dev:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 2}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=576289)
    ap.add_argument("--nnz", type=int, default=42_460_000)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--impl", default="float_pob", choices=["fixed", "float_pob", "float_stall"])
    ap.add_argument("--step-form", action="store_true", help="one GPU: hsb_iterate as a launch per step instead of one cooperative launch")
    ap.add_argument("--check", action="store_true", help="compare with a host iteration (oracle; small sizes)")
    ap.add_argument("--p2p", action="store_true",
                    help="N > 1: exchange the x blocks inside the update kernel over peer memory (hsb_axpb_to_peers) "
                         "instead of NCCL broadcasts issued by the host")
    args = ap.parse_args()
    from hisparse_b200 import capi, matgen, sharding
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rows, cols, indptr, indices, _ = matgen.rmat_csr(args.nodes, args.nnz, 0xC0FFEE04, symmetric=True, oversample=1.5)
    outdeg = np.maximum(np.bincount(indices, minlength=cols), 1)
    data = (1.0 / outdeg[indices]).astype(np.float32)
    r2, c2, ip2 = matgen.pad_csr(rows, cols, indptr, 128, 128)
    fixed = args.impl == "fixed"
    words = matgen.quantize_q824(data) if fixed else data.view(np.uint32)
    a_f, b_f = np.float32(0.85), np.float32(0.15)          # x is kept scaled by N so that Q8.24 has the range
    alpha = int(matgen.quantize_q824(a_f)) if fixed else int(a_f.view(np.uint32))
    beta = int(matgen.quantize_q824(b_f)) if fixed else int(b_f.view(np.uint32))
    x0f = np.ones(c2, np.float32)
    x0 = matgen.quantize_q824(x0f) if fixed else x0f.view(np.uint32)
    bounds = sharding.shard_bounds(ip2, world)
    sip, six, sw = sharding.extract_shard(ip2, indices, words, bounds[rank], bounds[rank + 1])
    ctx = capi.Context(local, args.impl)
    ctx.upload_matrix_csr(bounds[rank + 1] - bounds[rank], c2, sip, six, sw)
    ctx.upload_vector(x0)
    ctx.sync()
    if args.step_form:
        ctx.set_option("iterate_persistent", 0)
    p2p = args.p2p and dist is not None
    if p2p:
        mine = torch.from_numpy(ctx.peer_export()).cuda(local)
        blobs = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(blobs, mine)
        ctx.peer_connect(world, rank, torch.cat(blobs).cpu().numpy())
        dist.barrier()

    def iterate(n):
        if p2p:
            # compute + all-gather in one kernel, arrival flags polled by the next SpMV: the host only enqueues
            if not args.step_form:
                ctx.iterate_peers(n, alpha, beta, bounds[rank])      # the whole iteration loop: one resident kernel per GPU
                ctx.sync()
                return
            for _ in range(n):
                ctx.spmv()
                ctx.axpb_to_peers(alpha, beta, bounds[rank])
                ctx.vector_commit()
            ctx.sync()
            return
        if dist is None:
            # one GPU: hsb_iterate -- one cooperative launch for all n iterations (--step-form: a launch per step)
            ctx.iterate(n, alpha, beta)
            ctx.sync()
            return
        for _ in range(n):
            ctx.spmv()
            ctx.axpb_to_vector(alpha, beta, bounds[rank])
            if dist is not None:
                ctx.sync()                                  # the engine's stream -> torch's NCCL stream
                nxt = torch.as_tensor(_Dev(ctx.device_x_next(), c2), device="cuda:%d" % local)
                sharding.allgather_blocks(dist, nxt, bounds)
                torch.cuda.synchronize()
            ctx.vector_commit()
        ctx.sync()

    iterate(3)
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    iterate(args.iters)
    if dist is not None:
        dist.barrier()
    sec = (time.perf_counter() - t0) / args.iters
    ctx.spmv()
    y = ctx.download_result()
    out = {"nodes": r2, "nnz": int(ip2[-1]), "impl": args.impl, "n_gpus": world, "iters": args.iters,
           "ms_per_iteration": 1e3 * sec, "gops": 2.0 * int(ip2[-1]) / sec / 1e9,
           "exchange": ("peer-memory stores + arrival flags inside " + ("the update kernel" if args.step_form else "one resident kernel per GPU")) if p2p else
                       ("NCCL broadcasts issued by the host" if world > 1 else ("none (single GPU, launch per step)" if args.step_form else "none (single GPU, one cooperative launch)")),
           "what": "spmv + fused drain/axpb + exchange of the x blocks + commit, per iteration"}
    if args.check:
        from oracle import hsoracle
        port = hsoracle.Port()
        if fixed:
            x = x0.copy()
            for _ in range(args.iters + 3):
                x = hsoracle.axpb_q824(alpha, port.spmv_q824(ip2, indices, words, x), beta)
            want = port.spmv_q824(ip2, indices, words, x)[bounds[rank]:bounds[rank + 1]]
            ok = np.array_equal(y, want)
            if dist is not None:
                t = torch.tensor([int(ok)], device="cuda:%d" % local)
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
                ok = bool(t.item())
            out["parity"] = "bit-exact on every rank" if ok else "MISMATCH"
        else:
            x = x0f.astype(np.float64)
            for _ in range(args.iters + 3):
                y64, _ = port.spmv_f64(ip2, indices, data, x.astype(np.float32))
                x = 0.85 * y64 + 0.15
            y64, sa = port.spmv_f64(ip2, indices, data, x.astype(np.float32))
            y64, sa = y64[bounds[rank]:bounds[rank + 1]], sa[bounds[rank]:bounds[rank + 1]]
            err = float(np.max(np.abs(y.view(np.float32) - y64) / (sa + 1e-30)))
            out["parity"] = "max |err| / sum|a x| = %.2e" % err
    if rank == 0:
        print(json.dumps(out))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
