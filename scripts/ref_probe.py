"""Quick GPU probe of mode=reference: v5 kernel vs the 4-set kernel and the direct kernel on a few hundred rows per street, and
timings."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dask-holdem-simulator_b200"))
import holdem_b200 as hb

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

for nb in (3, 4, 5):
    d = torch.from_numpy(hb.random_situations(2048, seed=700 + nb, n_board=nb)).cuda()
    v5 = hb.hand_potential_counts(d, "reference", ref_variant=5)
    v4 = hb.hand_potential_counts(d, "reference", ref_variant=4)
    torch.cuda.synchronize()
    bad = torch.nonzero((v5 != v4).any(dim=1)).flatten()
    print(f"n_board {nb}: v5 == v4: {bad.numel() == 0} ({bad.numel()} rows differ)", flush=True)
    if bad.numel():
        i = int(bad[0])
        print("  sit", d[i].tolist(), "\n  v5", v5[i].tolist(), "\n  v4", v4[i].tolist(), flush=True)
    big = torch.from_numpy(hb.random_situations(65536, seed=800 + nb, n_board=nb)).cuda()
    t5 = timed(lambda: hb.hand_potential_counts(big, "reference", ref_variant=5))
    small = big[:8192].contiguous()
    t4 = timed(lambda: hb.hand_potential_counts(small, "reference", ref_variant=4), reps=1)
    ths = timed(lambda: hb.hand_strength_counts(big, "reference"))
    print(f"  v5 {65536 / t5 / 1e3:.2f} M situations/s, v4 {8192 / t4 / 1e3:.2f} M/s, HS {65536 / ths / 1e3:.1f} M/s", flush=True)
