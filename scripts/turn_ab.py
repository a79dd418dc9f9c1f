"""A/B of the turn kernels on the GPU: correctness against the oracle on a sample, then CUDA-event timings."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dask-holdem-simulator_b200"))
sys.path.insert(0, ROOT)
import holdem_b200 as hb  # noqa: E402
from oracle import oracle as O  # noqa: E402

hb._lib.init(0)
cases = [([0, 1], [2, 3, 4, 5]), ([13, 26], [2, 3, 4, 5]), ([12, 11], [2, 3, 4, 18]), ([25, 38], [2, 3, 4, 18]),
         ([0, 13], [2, 3, 15, 16]), ([12, 25], [38, 51, 0, 13]), ([39, 40], [41, 42, 43, 0]), ([39, 1], [41, 42, 43, 44]),
         ([50, 51], [49, 48, 47, 46]), ([14, 1], [31, 32, 33, 20])]
sit = np.concatenate([hb.pack_situations([c[0] for c in cases], [c[1] for c in cases]),
                      hb.random_situations(300, seed=5, n_board=4)])
want = O.hp_batch(sit, "treys")
d = torch.from_numpy(sit).cuda()
for v in (4, 3):
    got = hb.hand_potential_counts(d, "treys", turn_variant=v).cpu().numpy().astype(np.uint32)
    bad = np.nonzero((got != want).any(axis=1))[0]
    print(f"variant {v}: {len(bad)} mismatching rows of {len(sit)}", flush=True)
    for b in bad[:4]:
        print("  sit", sit[b].tolist(), "\n  got ", got[b].tolist(), "\n  want", want[b].tolist())
n = 1 << 20
big = torch.from_numpy(hb.random_situations(n, seed=20251022, n_board=4)).cuda()
ref = None
for v in (4, 3):
    out = torch.empty((n, 16), dtype=torch.int32, device="cuda")
    hb.hand_potential_counts(big, "treys", out=out, turn_variant=v)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        hb.hand_potential_counts(big, "treys", out=out, turn_variant=v)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"variant {v}: {ms:.3f} ms per {n} turn situations = {n / ms / 1e3:.2f} M situations/s", flush=True)
    if ref is None:
        ref = out.clone()
    else:
        print("   equal rows:", bool(torch.equal(ref, out)))
