"""torchrun --nproc-per-node N scripts/fused_allgather_check.py [--rows-per-rank R] [--time] [--out FILE]

Correctness of the multi-GPU path on REAL peer GPUs: the fused peer-store path (kernels write every rank's symmetric-memory
buffer over NVLink) and the NCCL all-gather path, each compared row by row -- on EVERY rank -- with a single-GPU run of the
whole batch (mixed streets, both modes, ragged N, both alternating buffers, a malformed row), plus the content hash of the
gathered buffer against the single-GPU hash.  Exit code 0 only if every rank agrees on everything; rank 0 prints one JSON
line (and writes it to --out).  tests/test_gpu_multi.py runs this under torchrun when the box has >= 2 GPUs.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dask-holdem-simulator_b200"))
import holdem_b200 as hb  # noqa: E402
from holdem_b200 import sharded  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows-per-rank", type=int, default=65536)
ap.add_argument("--time", action="store_true", help="also time both paths on 65,536 flop situations per rank")
ap.add_argument("--out", default=None)
args = ap.parse_args()

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
hb._lib.init(local)
n = args.rows_per_rank * world + 37
parts = [hb.random_situations(n // 2, seed=11), hb.random_situations(n - n // 2 - 100, seed=12, n_board=4),
         hb.random_situations(100, seed=13, n_board=5)]
sit_np = np.concatenate(parts)
sit_np = sit_np[np.random.default_rng(1).permutation(n)]
sit_np[5, 1] = sit_np[5, 0]                                  # one malformed row (rank 0's block)
sit = torch.from_numpy(sit_np).cuda()
checks = {}


def agree(name, ok):
    """AND over ranks; recorded once."""
    t = torch.tensor([1 if ok else 0], device="cuda", dtype=torch.int32)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    checks[name] = bool(int(t.item()))


for mode in ("treys", "reference"):
    sub = sit if mode == "treys" else sit[:min(n, 2048 * world + 5)].contiguous()
    ref = hb.hand_potential_counts(sub, mode)                                    # every rank computes everything itself
    ref_hash = hb.rows_hash(ref)
    a = hb.hand_potential_counts_sharded(sub, mode, fused=False).clone()
    agree(f"{mode}: nccl all-gather == single GPU (every rank, every row)", torch.equal(a, ref))
    agree(f"{mode}: hash(nccl all-gather) == hash(single GPU)", hb.rows_hash(a.contiguous()) == ref_hash)
    b = hb.hand_potential_counts_sharded(sub, mode, fused=True).clone()
    c = hb.hand_potential_counts_sharded(sub, mode, fused=True).clone()           # the second of the two alternating buffers
    agree(f"{mode}: fused peer stores == single GPU (every rank, every row)", torch.equal(b, ref))
    agree(f"{mode}: fused peer stores, second buffer == single GPU", torch.equal(c, ref))
    agree(f"{mode}: hash(fused) == hash(single GPU)", hb.rows_hash(b.contiguous()) == ref_hash)
    # shard hashes add up to the full hash
    lo, hi, per = hb.shard_bounds(sub.shape[0], world, rank)
    hv = hb.rows_hash(ref[lo:hi].contiguous(), first_row=lo) if hi > lo else 0
    part = torch.tensor([hv - (1 << 64) if hv >= (1 << 63) else hv], device="cuda", dtype=torch.int64)   # wraps like uint64
    dist.all_reduce(part, op=dist.ReduceOp.SUM)
    agree(f"{mode}: sum of shard hashes == full hash", (int(part.item()) & 0xFFFFFFFFFFFFFFFF) == ref_hash)
    agree(f"{mode}: malformed row marked on every rank", bool((b[5] == -1).all()) and bool((a[5] == -1).all()))

timing = {}
if args.time:
    flop = torch.from_numpy(hb.random_situations(65536 * world, seed=20251019)).cuda()
    for name, fused in (("nccl_allgather", False), ("fused_peer_stores", True)):
        for _ in range(5):
            hb.hand_potential_counts_sharded(flop, "treys", fused=fused)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            hb.hand_potential_counts_sharded(flop, "treys", fused=fused)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / 50], device="cuda", dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        timing[name] = {"ms_per_step": float(ms), "situations_per_sec": flop.shape[0] / float(ms) * 1e3}

ok = all(checks.values())
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    line = json.dumps({"what": "fused NVLink peer-store path and NCCL all-gather path vs a single-GPU run, every rank compares "
                               "every row", "n_gpus": world, "rows": n, "fused_path_error": sharded.fused_path_error(),
                       "checks": checks, "timing": timing, "all_ok": ok})
    print(line, flush=True)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            f.write(line + "\n")
sys.exit(0 if ok else 1)
