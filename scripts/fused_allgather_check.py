"""torchrun --nproc-per-node N scripts/fused_allgather_check.py: the fused peer-store path (kernels write every rank's
symmetric-memory buffer over NVLink) against the NCCL all-gather path and a single-GPU run; timings of both."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dask-holdem-simulator_b200"))
import holdem_b200 as hb  # noqa: E402
from holdem_b200 import sharded  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.pop("NCCL_DEBUG", None)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
hb._lib.init(local)
n = 65536 * world + 37
parts = [hb.random_situations(n // 2, seed=11), hb.random_situations(n - n // 2 - 100, seed=12, n_board=4),
         hb.random_situations(100, seed=13, n_board=5)]
sit_np = np.concatenate(parts)
sit_np = sit_np[np.random.default_rng(1).permutation(n)]
sit = torch.from_numpy(sit_np).cuda()
ok = True
for mode in ("treys", "reference") if n < 400000 else ("treys",):
    sub = sit if mode == "treys" else sit[:4096 * world + 5]
    ref = hb.hand_potential_counts(sub, mode)                                    # every rank computes everything
    a = hb.hand_potential_counts_sharded(sub, mode, fused=False).clone()
    b = hb.hand_potential_counts_sharded(sub, mode, fused=True).clone()
    c = hb.hand_potential_counts_sharded(sub, mode, fused=True).clone()           # second buffer
    ok = ok and bool(torch.equal(a, ref)) and bool(torch.equal(b, ref)) and bool(torch.equal(c, ref))
    if rank == 0:
        print(mode, "all-gather == single:", bool(torch.equal(a, ref)), " fused == single:", bool(torch.equal(b, ref)),
              bool(torch.equal(c, ref)), "symm error:", sharded._SYMM_ERROR, flush=True)
flop = torch.from_numpy(hb.random_situations(65536 * world, seed=20251019)).cuda()
for name, fused in (("nccl all-gather", False), ("fused peer stores", True)):
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
    if rank == 0:
        print(f"{name}: {float(ms):.3f} ms per {flop.shape[0]} situations = {flop.shape[0] / float(ms) / 1e3:.2f} M situations/s",
              flush=True)
t = torch.tensor([int(ok)], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("ALL OK" if int(t) else "MISMATCH", flush=True)
dist.barrier()
dist.destroy_process_group()
