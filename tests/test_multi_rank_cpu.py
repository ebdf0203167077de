"""CPU, world_size 2, gloo: the N>1 path's host logic -- nnz-balanced row-block shards, broadcast
of x from rank 0, per-rank SpMV of the shard, gather of the y blocks -- reproduces the
single-process result bit for bit (fixed point). The per-shard SpMV is done by the oracle here
(no GPU in this container); on the GPU box bench.py --gpus N runs the same plumbing over NCCL with
the CUDA engine."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port_no, out_dir):
    sys.path.insert(0, ROOT)
    from hisparse_b200 import matgen, sharding
    from oracle import hsoracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    port = hsoracle.Port()
    # every rank derives the same matrix from the seed (as bench.py does); only rank 0 owns x
    rows, cols, indptr, indices, data = matgen.rmat_csr(6000, 90000, 77)
    r2, c2, ip2 = matgen.pad_csr(rows, cols, indptr, 128, 8)
    words = port.quantize(data)
    bounds = sharding.shard_bounds(ip2, world)
    r0, r1 = bounds[rank], bounds[rank + 1]
    sip, six, sw = sharding.extract_shard(ip2, indices, words, r0, r1)
    x = torch.zeros(c2, dtype=torch.int64)
    if rank == 0:
        xf = np.zeros(c2, np.float32)
        xf[:cols] = np.random.default_rng(5).random(cols, dtype=np.float32)
        x = torch.from_numpy(port.quantize(xf).astype(np.int64))
    dist.broadcast(x, 0)                                    # the ncclBroadcast(x) of the GPU path
    xw = x.numpy().astype(np.uint32)
    y_block = port.spmv_q824(sip, six, sw, xw)              # stand-in for hsb_spmv on this rank's shard
    # gather of unequal blocks on rank 0 (padded all_gather, as NCCL has no gatherv)
    longest = max(sharding.gather_counts(bounds))
    pad = torch.zeros(longest, dtype=torch.int64)
    pad[: y_block.size] = torch.from_numpy(y_block.astype(np.int64))
    got = [torch.zeros(longest, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(got, pad)
    if rank == 0:
        y = np.concatenate([got[g][: bounds[g + 1] - bounds[g]].numpy() for g in range(world)]).astype(np.uint32)
        want = port.spmv_q824(ip2, indices, words, xw)
        np.save(os.path.join(out_dir, "ok.npy"), np.array([int(np.array_equal(y, want)), bounds[1]]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_row_block_sharding(tmp_path):
    world = 2
    port_no = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port_no, str(tmp_path)), nprocs=world, join=True)
    ok, split = np.load(tmp_path / "ok.npy")
    assert ok == 1
    assert 0 < split < 6016 and split % 128 == 0


def test_shard_bounds_balance():
    sys.path.insert(0, ROOT)
    from hisparse_b200 import matgen, sharding
    rows, cols, indptr, indices, data = matgen.rmat_csr(20000, 400000, 3)
    r2, c2, ip2 = matgen.pad_csr(rows, cols, indptr, 128, 8)
    for world in (1, 2, 4, 8):
        b = sharding.shard_bounds(ip2, world)
        assert b[0] == 0 and b[-1] == r2 and len(b) == world + 1 and all(x <= y for x, y in zip(b, b[1:]))
        nnz = [int(ip2[b[g + 1]]) - int(ip2[b[g]]) for g in range(world)]
        assert max(nnz) <= 1.25 * (sum(nnz) / world) + 128 * 500
        assert all(x % 128 == 0 for x in b[:-1])


def _pagerank_worker(rank, world, port_no, out_dir):
    """x <- alpha (*) A x (+) beta over row-block shards: every rank multiplies its shard, writes its slice
    of the next vector at its row offset and the slices are all-gathered in place -- the loop of
    tests/pagerank.py with the oracle standing in for the GPU engine."""
    sys.path.insert(0, ROOT)
    from hisparse_b200 import matgen, sharding
    from oracle import hsoracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    port = hsoracle.Port()
    rows, cols, indptr, indices, data = matgen.rmat_csr(4000, 60000, 91)
    r2, c2, ip2 = matgen.pad_csr(rows, cols, indptr, 128, 128)
    words = port.quantize((data * np.float32(0.05)).astype(np.float32))
    alpha, beta = int(port.quantize(np.float32([0.85]))[0]), int(port.quantize(np.float32([0.01]))[0])
    bounds = sharding.shard_bounds(ip2, world)
    sip, six, sw = sharding.extract_shard(ip2, indices, words, bounds[rank], bounds[rank + 1])
    x = port.quantize(np.full(c2, 0.25, np.float32))
    for _ in range(5):
        y_block = port.spmv_q824(sip, six, sw, x)                         # hsb_spmv on this rank's shard
        nxt = torch.zeros(c2, dtype=torch.int64)
        nxt[bounds[rank]:bounds[rank + 1]] = torch.from_numpy(            # hsb_axpb_to_vector(alpha, beta, bounds[rank])
            hsoracle.axpb_q824(alpha, y_block, beta).astype(np.int64))
        sharding.allgather_blocks(dist, nxt, bounds)
        x = nxt.numpy().astype(np.uint32)                                 # hsb_vector_commit
    if rank == 0:
        ref = port.quantize(np.full(c2, 0.25, np.float32))
        for _ in range(5):
            ref = hsoracle.axpb_q824(alpha, port.spmv_q824(ip2, indices, words, ref), beta)
        np.save(os.path.join(out_dir, "pr.npy"), np.array([int(np.array_equal(x, ref))]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_iteration(tmp_path):
    world = 2
    port_no = 31500 + os.getpid() % 2000
    mp.spawn(_pagerank_worker, args=(world, port_no, str(tmp_path)), nprocs=world, join=True)
    assert np.load(tmp_path / "pr.npy")[0] == 1
