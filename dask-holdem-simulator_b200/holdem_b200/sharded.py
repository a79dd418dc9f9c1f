"""Multi-GPU sharding of a batch of independent situations.

Situations are independent, so the batch is split into contiguous blocks, one per rank (one process per GPU); the
tables are replicated.  Each rank's kernel writes straight into its own slice of the full [N_padded,16] result
buffer and one in-place all-gather (NCCL over NVLink/NVSwitch on GPUs; gloo in the CPU tests of the host logic)
makes every rank hold all rows.  There is no other data-path collective.

Fused path (CUDA, one node): the result buffer is symmetric memory (torch.distributed._symmetric_memory, peer-mapped over
NVLink/NVSwitch) and the HP kernels store every result row into all ranks' buffers themselves -- compute and all-gather are
one launch, the transfer overlaps the computation row by row (holdem_hp_batch_peers).  One symmetric-memory barrier per call
orders the ranks (two buffers alternate, so the barrier of the previous call also protects the buffer being overwritten);
NCCL is not involved in the data path.

Which path a (group, device) uses is decided COLLECTIVELY, once: every rank tries the local part of the symmetric-memory
set-up, the ranks all-reduce (MIN) an "ok" flag, and only if everyone succeeded do they rendezvous; a second all-reduce
after the rendezvous confirms it.  No rank can therefore end up in the all-gather while its peers wait in the
symmetric-memory barrier.  Errors raised by the kernels' entry point (bad arguments, CUDA errors) do not change the path:
the failing rank still reaches the barrier and raises afterwards.
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


class _Fused:
    """Per (group, device) state of the fused path: capability (decided collectively) and ONE pair of symmetric buffers that
    grows to the largest batch seen (capacity in rows, per rank block x world) instead of one pair per batch size."""

    def __init__(self):
        self.available = None       # None = not decided yet
        self.reason = None
        self.capacity = 0           # rows of each buffer
        self.bufs = []
        self.hdls = []
        self.next = 0


_FUSED = {}                         # (device index, id(group)) -> _Fused


def fused_state(device, group=None) -> _Fused:
    return _FUSED.setdefault((device.index, id(group)), _Fused())


def fused_path_error(device=None, group=None):
    """None while the fused path is in use (or undecided); otherwise why this (group, device) uses the all-gather."""
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    st = _FUSED.get((device.index, id(group)))
    return None if st is None or st.available is not False else st.reason


def _all_ok(ok: bool, device, group) -> bool:
    t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return bool(int(t.item()))


def _ensure_buffers(st: _Fused, rows: int, device, group) -> bool:
    """Collective.  Makes st hold two symmetric [capacity >= rows, 16] buffers; returns False (on every rank) if that is
    impossible anywhere."""
    if st.available is False:
        return False
    if st.available and st.capacity >= rows:
        return True
    # (re)allocation: every rank computes the same `rows`, so every rank takes this branch in the same call
    err = None
    bufs = []
    try:
        import torch.distributed._symmetric_memory as symm_mem
        cap = max(rows, 2 * st.capacity)
        for _ in range(2):
            bufs.append(symm_mem.empty((cap, HP_ROW), dtype=torch.int32, device=device))
    except Exception as e:          # symmetric memory not supported here (set-up only: nothing else is caught)
        err = repr(e)
    if not _all_ok(err is None, device, group):
        st.available, st.reason = False, err or "symmetric memory unavailable on another rank"
        return False
    hdls = []
    try:
        g = group if group is not None else dist.group.WORLD
        for t in bufs:
            hdls.append(symm_mem.rendezvous(t, g))
    except Exception as e:
        err = repr(e)
    if not _all_ok(err is None, device, group):
        st.available, st.reason = False, err or "symmetric-memory rendezvous failed on another rank"
        return False
    st.bufs, st.hdls, st.capacity, st.next, st.available = bufs, hdls, cap, 0, True
    return True


def _fused_peer_store(st: _Fused, sit, mode, lo, hi, per, world, rank):
    """Kernels store their rows into every rank's symmetric buffer; returns the buffer (first per*world rows valid) of this
    call.  The returned tensor is valid until the second next call on this (group, device): two buffers alternate."""
    from . import _ops
    from .api import mode_id
    k = st.next
    st.next = k ^ 1
    buf, hdl = st.bufs[k], st.hdls[k]
    # No barrier is needed before overwriting: this buffer was last used two calls ago, and every rank has since passed the
    # barrier that ends the previous call -- which its stream reaches only after everything it enqueued on the result of
    # that older call (consumers on other streams must order themselves against this one).
    n_local = hi - lo
    failure = None
    if n_local > 0:
        try:
            block = sit[lo:hi].contiguous()
            row_bytes = HP_ROW * 4
            ptrs = [int(p) + rank * per * row_bytes for r, p in enumerate(hdl.buffer_ptrs) if r != rank]
            mine = buf[rank * per: rank * per + n_local]
            _ops.load().hp_peers(block, mode_id(mode), mine, ptrs)
        except Exception as e:      # keep the ranks in step: reach the barrier first, raise afterwards
            failure = e
    hdl.barrier(channel=1)          # every rank's stores have landed (stream-ordered behind the kernels)
    if failure is not None:
        raise failure
    return buf


def hand_potential_counts_sharded(sit: torch.Tensor, mode="treys", group=None, compute=None, gather=True, fused=None):
    """sit: the FULL uint8 [N,8] batch, present on every rank (or at least this rank's block valid).
    Returns int32 [N,16] with all rows on every rank (gather=True) or only this rank's block filled.
    `compute(sit_block, mode, out_block)` defaults to the CUDA kernel; the gloo CPU tests inject a stand-in.
    fused: True = kernels store into all ranks' symmetric-memory buffers (no NCCL call; the result is valid until the second
    next call on this group and device; raises on every rank if the path is unavailable), False = in-place NCCL all-gather,
    None = fused when available (decided collectively at first use)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = sit.shape[0]
    lo, hi, per = shard_bounds(n, world, rank)
    if compute is None and gather and world > 1 and sit.is_cuda and fused is not False:
        st = fused_state(sit.device, group)
        if _ensure_buffers(st, per * world, sit.device, group):
            return _fused_peer_store(st, sit, mode, lo, hi, per, world, rank)[:n]
        if fused:
            raise RuntimeError(f"fused peer-store path unavailable: {st.reason}")
    if compute is None:
        from .api import hand_potential_counts

        def compute(block, mode_, out_block):
            hand_potential_counts(block, mode_, out=out_block)
    full = torch.zeros((per * world, HP_ROW), dtype=torch.int32, device=sit.device)
    mine = full[rank * per: rank * per + (hi - lo)]
    if hi > lo:
        compute(sit[lo:hi].contiguous(), mode, mine)
    if gather and world > 1:
        # in place: this rank's input is its own slice of the output buffer
        dist.all_gather_into_tensor(full, full[rank * per:(rank + 1) * per], group=group)
    return full[:n]
