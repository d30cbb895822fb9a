"""Batch sharding across the GPUs of one box.

Every grid of a batch is independent end to end (all reductions of the path are per sample and
BatchNorm runs in inference mode), so the path shards with no data-path collective: each rank
(one process per GPU) takes a contiguous slice of the batch and replicated weights.  The only
exchange is a final all-reduce of a few floats of error statistics (sum of squared errors, sum of
squared norms, max, count) -- torch.distributed over NCCL on the GPUs, gloo in the CPU tests.
The reference has no inference parallelism; this is new.
"""
import os

import torch
import torch.distributed as dist


def shard_bounds(n, world_size, rank):
    """Contiguous [lo, hi) slice of n items for `rank`: sizes differ by at most one (numpy.array_split)."""
    per, extra = divmod(int(n), int(world_size))
    lo = rank * per + min(rank, extra)
    return lo, lo + per + (1 if rank < extra else 0)


def shard_problem(problem, world_size, rank):
    """Slice every [B, ...] tensor of a problem dict to this rank's samples."""
    B = next(iter(problem.values())).shape[0]
    lo, hi = shard_bounds(B, world_size, rank)
    return {k: v[lo:hi] for k, v in problem.items()}


def bucket_by_shape(shapes):
    """Group sample indices by grid shape (config 3: one forward per distinct (nx, ny))."""
    buckets = {}
    for i, s in enumerate(shapes):
        buckets.setdefault(tuple(s), []).append(i)
    return buckets


def init_from_env(backend=None):
    """Join the process group described by RANK/WORLD_SIZE/MASTER_* (torchrun); no-op for 1 process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


class ErrorStats:
    """Accumulates sum|a-b|^2, sum|b|^2, max|a-b| and the sample count; combine() all-reduces them."""

    def __init__(self, device="cpu"):
        self.sums = torch.zeros(3, dtype=torch.float64, device=device)   # sq_err, sq_ref, count
        self.maxv = torch.zeros(1, dtype=torch.float64, device=device)

    def update(self, a, b):
        d = (a.double() - b.double())
        self.sums += torch.stack([d.pow(2).sum(), b.double().pow(2).sum(),
                                  torch.tensor(float(a.shape[0]), dtype=torch.float64, device=d.device)])
        self.maxv = torch.maximum(self.maxv, d.abs().max().reshape(1))
        return self

    def combine(self):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.sums, op=dist.ReduceOp.SUM)
            dist.all_reduce(self.maxv, op=dist.ReduceOp.MAX)
        return self

    def result(self):
        s = self.sums.cpu()
        return {"rel_l2": float((s[0] / s[1]).sqrt()) if float(s[1]) > 0 else float("nan"),
                "max_abs_err": float(self.maxv.cpu()), "samples": int(s[2])}
