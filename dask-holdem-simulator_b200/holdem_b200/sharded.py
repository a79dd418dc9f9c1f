"""Multi-GPU sharding of a batch of independent situations.

Situations are independent, so the batch is split into contiguous blocks, one per rank (one process per GPU); the
tables are replicated.  Each rank's kernel writes straight into its own slice of the full [N_padded,16] result
buffer and one in-place all-gather (NCCL over NVLink/NVSwitch on GPUs; gloo in the CPU tests of the host logic)
makes every rank hold all rows.  There is no other data-path collective.
"""
import torch
import torch.distributed as dist

from ._lib import HP_ROW


def shard_bounds(n: int, world_size: int, rank: int):
    """Contiguous block split: rank g owns rows [g*per, min(n,(g+1)*per)), per = ceil(n / world_size)."""
    per = (n + world_size - 1) // world_size
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi, per


def hand_potential_counts_sharded(sit: torch.Tensor, mode="treys", group=None, compute=None, gather=True):
    """sit: the FULL uint8 [N,8] batch, present on every rank (or at least this rank's block valid).
    Returns int32 [N,16] with all rows on every rank (gather=True) or only this rank's block filled.
    `compute(sit_block, mode, out_block)` defaults to the CUDA kernel; the gloo CPU tests inject a stand-in."""
    if compute is None:
        from .api import hand_potential_counts

        def compute(block, mode_, out_block):
            hand_potential_counts(block, mode_, out=out_block)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = sit.shape[0]
    lo, hi, per = shard_bounds(n, world, rank)
    full = torch.zeros((per * world, HP_ROW), dtype=torch.int32, device=sit.device)
    mine = full[rank * per: rank * per + (hi - lo)]
    if hi > lo:
        compute(sit[lo:hi].contiguous(), mode, mine)
    if gather and world > 1:
        # in place: this rank's input is its own slice of the output buffer
        dist.all_gather_into_tensor(full, full[rank * per:(rank + 1) * per], group=group)
    return full[:n]
