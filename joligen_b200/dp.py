"""Data-parallel plumbing (one process per GPU, torch.distributed).

The path shards by SAMPLE: every rank runs the full step on its own batch shard; the only exchange is
the gradient all-reduce before the optimiser (DDP semantics, base_model.py:725-737: mean over ranks).
Here the gradient is ONE flat fp32 buffer, so the exchange is a single SUM all-reduce (NCCL over
NVLink/NVSwitch on the GPU box, gloo in the CPU tests) and the 1/world factor is folded into the
fused optimizer kernel.
"""
import torch
import torch.distributed as dist


def world_size(pg=None):
    return dist.get_world_size(pg) if (dist.is_available() and dist.is_initialized()) else 1


def allreduce_sum_(flat_grad: torch.Tensor, pg=None):
    """In-place SUM all-reduce of the flat gradient buffer; returns the scale (1/world) to apply."""
    w = world_size(pg)
    if w > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=pg)
    return 1.0 / w


def broadcast_(flat_params: torch.Tensor, pg=None, src=0):
    if world_size(pg) > 1:
        dist.broadcast(flat_params, src=src, group=pg)


def shard_seed(base_seed: int, rank: int) -> int:
    """Per-rank synthetic-data seed (seeds 1234 + rank)."""
    return base_seed + rank
