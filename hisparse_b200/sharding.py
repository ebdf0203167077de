"""Row-block sharding of a CSR across GPUs: rank g holds rows
[bounds[g], bounds[g+1]) -- contiguous blocks balanced by non-zeros, boundaries rounded to the
reference's row granularity (128 rows: PACK_SIZE * NUM_HBM_CHANNELS, sw/data_formatter.h:475) --
plus a full replica of x, and produces y for its block. The only exchange steps are a broadcast of
x before the (repeated) SpMV and, when a caller needs the whole y on one rank, a gather of the
blocks. Nothing here touches the GPU; bench.py and the tests drive it with torch.distributed."""
import numpy as np


def shard_bounds(indptr, world, granularity=128):
    """nnz-balanced row boundaries, multiples of `granularity` (last one = rows)."""
    indptr = np.asarray(indptr, dtype=np.int64)
    rows = indptr.size - 1
    nnz = int(indptr[-1])
    bounds = [0]
    for g in range(1, world):
        target = nnz * g // world
        r = int(np.searchsorted(indptr, target, side="left"))
        r = min(rows, (r + granularity // 2) // granularity * granularity)
        bounds.append(max(bounds[-1], r))
    bounds.append(rows)
    return bounds


def extract_shard(indptr, indices, data, r0, r1):
    """CSR of rows [r0, r1) with indptr rebased to 0."""
    e0, e1 = int(indptr[r0]), int(indptr[r1])
    return ((indptr[r0:r1 + 1].astype(np.int64) - e0).astype(np.uint32), indices[e0:e1], data[e0:e1])


def gather_counts(bounds):
    return [bounds[g + 1] - bounds[g] for g in range(len(bounds) - 1)]


def allgather_blocks(dist, x_full, bounds):
    """In-place all-gather of unequal contiguous blocks: rank g owns x_full[bounds[g]:bounds[g+1]] (a torch
    tensor on the rank's device, or on the CPU with gloo) and receives everybody else's block. NCCL has no
    all-gather-v; with a handful of ranks one broadcast per block is the simplest exact equivalent and
    moves the same bytes over NVLink. Used by iterative callers: x(k+1) = f(y(k)) where every rank
    produced the y block of its row shard (hsb_axpb_to_vector writes it at its row offset)."""
    for g in range(len(bounds) - 1):
        if bounds[g + 1] > bounds[g]:
            dist.broadcast(x_full[bounds[g]:bounds[g + 1]], src=g)

