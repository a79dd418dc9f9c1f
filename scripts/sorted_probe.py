"""Which work order suits the flop kernel best?  Times the kernel WITHOUT its own ordering pass (flop_variant=6) on the bench rows
pre-sorted on the host by several keys, and the production kernel (its own texture ordering) on the unsorted rows."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dask-holdem-simulator_b200"))
import holdem_b200 as hb

def timed(fn, reps=20):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

n = 65536
sit = hb.random_situations(n, seed=20251019)
bs = sit[:, 2:5] // 13
hs = sit[:, 0:5] // 13
nsuit = np.array([len(set(r)) for r in bs.tolist()])                       # 1 monotone, 2 two-tone, 3 rainbow
ourmax = np.array([max(np.bincount(r, minlength=4)) for r in hs])        # cards of our best suit among hole + board
suited_hole = (sit[:, 0] // 13 == sit[:, 1] // 13).astype(int)
brk = np.sort(sit[:, 2:5] % 13, axis=1)
paired = ((brk[:, 0] == brk[:, 1]) | (brk[:, 1] == brk[:, 2])).astype(int)
keys = {
    "input order": np.arange(n),
    "texture": nsuit,
    "texture, our flush draw": nsuit * 10 + (ourmax >= 3) * 1,
    "texture, our flush draw size": nsuit * 10 + (5 - ourmax),
    "texture, paired board": nsuit * 10 + paired,
    "texture, draw size, paired": nsuit * 100 + (5 - ourmax) * 10 + paired,
}
out = torch.empty((n, 16), dtype=torch.int32, device="cuda")
for name, key in keys.items():
    order = np.argsort(key, kind="stable")
    d = torch.from_numpy(np.ascontiguousarray(sit[order])).cuda()
    ms = timed(lambda: hb.hand_potential_counts(d, "treys", out=out, flop_variant=6))
    print(f"{name:32s}: {ms:.4f} ms = {n / ms / 1e3:.2f} M situations/s", flush=True)
d = torch.from_numpy(sit).cuda()
ms = timed(lambda: hb.hand_potential_counts(d, "treys", out=out))
print(f"{'production (own ordering pass)':32s}: {ms:.4f} ms = {n / ms / 1e3:.2f} M situations/s", flush=True)
