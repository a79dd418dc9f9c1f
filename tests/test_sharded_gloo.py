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
