"""Rank plumbing of the multi-GPU runs (one process per GPU, torch.distributed).

Round-1 status of the element partition: not implemented -- N ranks simulate N independent sheets
("replicas", weak scaling, no data-path collective).  What is shared across ranks is only the measurement protocol:
barrier, device-side timing, MAX over ranks of the elapsed time, SUM of the units processed."""
import torch
import torch.distributed as dist


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def barrier():
    if world() > 1:
        dist.barrier()


def aggregate(units, elapsed_ms, device="cpu"):
    """whole-job throughput inputs: (sum over ranks of units, max over ranks of each elapsed time in `elapsed_ms`)"""
    if world() == 1:
        return float(units), [float(t) for t in elapsed_ms]
    t = torch.tensor([float(x) for x in elapsed_ms], dtype=torch.float64, device=device)
    u = torch.tensor([float(units)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(u[0]), [float(x) for x in t]


def shard_sheets(n_sheets, rank, world_size):
    """independent sheets (objects) are dealt round-robin to ranks; returns the ids this rank owns"""
    return list(range(rank, n_sheets, world_size))
