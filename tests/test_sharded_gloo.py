"""world_size-2 gloo test of the sharding host logic on CPU: block split, in-place all-gather, padding when N is not
a multiple of the world size.  The compute step is injected (the CPU oracle stands in for the CUDA kernel: here the
oracle is the checker's own stand-in, not a product path)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, ret):
    import sys
    from conftest import PKG, ROOT  # noqa: F401  (sets sys.path)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import holdem_b200 as hb
    from oracle import oracle as O
    sit_np = hb.random_situations(n, seed=5, n_board=5)  # river boards: HS only, fast on CPU
    calls = []

    def compute(block, mode, out_block):
        calls.append(block.shape[0])
        out_block.copy_(torch.from_numpy(O.hp_batch(block.numpy(), mode).astype(np.int32)))

    full = hb.hand_potential_counts_sharded(torch.from_numpy(sit_np), "treys", compute=compute)
    lo, hi, per = hb.shard_bounds(n, world, rank)
    ok = full.shape == (n, 16) and calls == ([hi - lo] if hi > lo else [])
    want = O.hp_batch(sit_np, "treys").astype(np.int32)
    ok = ok and np.array_equal(full.numpy(), want)
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def _run(n):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, n, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_sharded_even():
    _run(64)


def test_sharded_ragged():
    _run(37)


def test_sharded_single_process_no_group():
    import holdem_b200 as hb
    from oracle import oracle as O
    sit_np = hb.random_situations(5, seed=9, n_board=5)

    def compute(block, mode, out_block):
        out_block.copy_(torch.from_numpy(O.hp_batch(block.numpy(), mode).astype(np.int32)))

    full = hb.hand_potential_counts_sharded(torch.from_numpy(sit_np), "reference", compute=compute)
    assert np.array_equal(full.numpy(), O.hp_batch(sit_np, "reference").astype(np.int32))


class _FakeHandle:
    def __init__(self, world):
        self.buffer_ptrs = [0] * world

    def barrier(self, channel=0):
        dist.barrier()


def _agreement_worker(rank, world, port, fail_rank, fail_stage, ret):
    """The fused-vs-all-gather decision is collective: symmetric memory is simulated by a fake module that fails on ONE rank (at
    allocation or at rendezvous); every rank must come to the same answer -- nobody may end up in the all-gather while its peer
    waits in the symmetric-memory barrier (round-1 advisor finding)."""
    import sys
    import types
    from conftest import PKG, ROOT  # noqa: F401
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fake = types.ModuleType("torch.distributed._symmetric_memory")

    def empty(shape, dtype=None, device=None):
        if rank == fail_rank and fail_stage == "alloc":
            raise RuntimeError("simulated: symmetric memory not supported on this rank")
        return torch.zeros(shape, dtype=dtype)

    def rendezvous(t, group):
        if rank == fail_rank and fail_stage == "rendezvous":
            raise RuntimeError("simulated: rendezvous failed on this rank")
        return _FakeHandle(world)

    fake.empty, fake.rendezvous = empty, rendezvous
    sys.modules["torch.distributed._symmetric_memory"] = fake
    import torch.distributed as d
    d._symmetric_memory = fake
    from holdem_b200 import sharded
    dev = torch.device("cpu")
    st = sharded.fused_state(dev, None)
    ok1 = sharded._ensure_buffers(st, 64, dev, None)
    ok2 = sharded._ensure_buffers(st, 128, dev, None)            # a second call does not flip the decision
    expect = fail_rank is None
    good = ok1 == expect and ok2 == expect and (st.available is expect)
    if expect:
        good = good and st.capacity >= 128 and len(st.bufs) == 2
    else:
        good = good and st.reason is not None and sharded.fused_path_error(dev) is not None
    ret[rank] = bool(good)
    dist.barrier()
    dist.destroy_process_group()


def _run_agreement(fail_rank, fail_stage):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_agreement_worker, args=(world, _free_port(), fail_rank, fail_stage, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_fused_path_decision_is_collective():
    _run_agreement(None, None)                 # everyone succeeds: fused path, one growing pair of buffers
    _run_agreement(1, "alloc")                 # one rank cannot allocate: both ranks use the all-gather
    _run_agreement(0, "rendezvous")            # one rank fails at the rendezvous: both ranks use the all-gather
