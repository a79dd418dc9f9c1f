"""Multi-GPU sharding of a batch of independent situations.

Situations are independent, so the batch is split into contiguous blocks, one per rank (one process per GPU); the
tables are replicated.  Each rank's kernel writes straight into its own slice of the full [N_padded,16] result
buffer and one in-place all-gather (NCCL over NVLink/NVSwitch on GPUs; gloo in the CPU tests of the host logic)
makes every rank hold all rows.  There is no other data-path collective.

Fused path (CUDA, one node): the result buffer is symmetric memory (torch.distributed._symmetric_memory, peer-mapped over
NVLink/NVSwitch) and the HP kernels store every result row into all ranks' buffers themselves -- compute and all-gather are
one launch, the transfer overlaps the computation row by row (holdem_hp_batch_peers).  One symmetric-memory barrier per call
orders the ranks (two buffers alternate, so the barrier of the previous call also protects the buffer being overwritten);
NCCL is not involved.  `fused=None` tries this path and falls back to the all-gather when symmetric
memory is unavailable.
"""
import ctypes
import torch
import torch.distributed as dist

from ._lib import HP_ROW


def shard_bounds(n: int, world_size: int, rank: int):
    """Contiguous block split: rank g owns rows [g*per, min(n,(g+1)*per)), per = ceil(n / world_size)."""
    per = (n + world_size - 1) // world_size
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi, per


_SYMM = {}          # (rows, device index, group id) -> [handle, buffers (two, used alternately), next]
_SYMM_ERROR = None


def _symmetric_buffers(rows, device, group):
    """Two peer-mapped [rows,16] int32 buffers (alternated between calls) and their rendezvous handles."""
    import torch.distributed._symmetric_memory as symm_mem
    key = (rows, device.index, id(group))
    ent = _SYMM.get(key)
    if ent is None:
        g = group if group is not None else dist.group.WORLD
        bufs, hdls = [], []
        for _ in range(2):
            t = symm_mem.empty((rows, HP_ROW), dtype=torch.int32, device=device)
            hdls.append(symm_mem.rendezvous(t, g))
            bufs.append(t)
        ent = {"bufs": bufs, "hdls": hdls, "next": 0}
        _SYMM[key] = ent
    return ent


def _fused_peer_store(sit, mode, group, lo, hi, per, world, rank):
    """Kernels store their rows into every rank's symmetric buffer; returns the buffer ([per*world,16]) of this call.
    The returned tensor is valid until the second next call with the same shape (two buffers alternate)."""
    from . import _lib
    from .api import _enter, mode_id
    ent = _symmetric_buffers(per * world, sit.device, group)
    k = ent["next"]
    ent["next"] = k ^ 1
    buf, hdl = ent["bufs"][k], ent["hdls"][k]
    # No barrier is needed before overwriting: this buffer was last used two calls ago, and every rank has since passed the
    # barrier that ends the previous call -- which its stream reaches only after everything it enqueued on the result of
    # that older call (consumers on other streams must order themselves against this one).
    n_local = hi - lo
    if n_local > 0:
        block = sit[lo:hi].contiguous()
        lib, st = _enter(block)
        row_bytes = HP_ROW * 4
        ptrs = [int(p) + rank * per * row_bytes for r, p in enumerate(hdl.buffer_ptrs) if r != rank]
        arr = (ctypes.c_void_p * max(1, len(ptrs)))(*ptrs)
        mine = buf[rank * per: rank * per + n_local]
        _lib.check(lib.holdem_hp_batch_peers(block.data_ptr(), n_local, mode_id(mode), mine.data_ptr(), arr, len(ptrs), st))
    hdl.barrier(channel=1)                       # every rank's stores have landed (stream-ordered behind the kernels)
    return buf


def hand_potential_counts_sharded(sit: torch.Tensor, mode="treys", group=None, compute=None, gather=True, fused=None):
    """sit: the FULL uint8 [N,8] batch, present on every rank (or at least this rank's block valid).
    Returns int32 [N,16] with all rows on every rank (gather=True) or only this rank's block filled.
    `compute(sit_block, mode, out_block)` defaults to the CUDA kernel; the gloo CPU tests inject a stand-in.
    fused: True = kernels store into all ranks' symmetric-memory buffers (no NCCL call; the result is valid until the second
    next call of the same shape), False = in-place NCCL all-gather, None = fused when available."""
    global _SYMM_ERROR
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if compute is None and gather and world > 1 and sit.is_cuda and fused is not False and (fused or _SYMM_ERROR is None):
        lo, hi, per = shard_bounds(sit.shape[0], world, rank)
        try:
            return _fused_peer_store(sit, mode, group, lo, hi, per, world, rank)[:sit.shape[0]]
        except Exception as e:  # symmetric memory unavailable on this system: remember and use the collective
            if fused:
                raise
            _SYMM_ERROR = repr(e)
    if compute is None:
        from .api import hand_potential_counts

        def compute(block, mode_, out_block):
            hand_potential_counts(block, mode_, out=out_block)
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
