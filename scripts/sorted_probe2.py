"""Host-sorted work order probes for the reference-mode kernel (all streets) and the Treys-mode turn kernel."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dask-holdem-simulator_b200"))
import holdem_b200 as hb

def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

n = 65536
for nb in (3, 4, 5):
    sit = hb.random_situations(n, seed=20251019 + nb, n_board=nb)
    bs = sit[:, 2:2 + nb] // 13
    maxb = np.array([max(np.bincount(r, minlength=4)) for r in bs])          # board cards of the most frequent suit
    for name, key in (("input order", np.arange(n)), ("most-suited boards first", -maxb)):
        order = np.argsort(key, kind="stable")
        d = torch.from_numpy(np.ascontiguousarray(sit[order])).cuda()
        ms = timed(lambda: hb.hand_potential_counts(d, "reference"))
        line = f"n_board {nb} reference {name:26s}: {n / ms / 1e3:.2f} M/s"
        if nb == 4:
            big = torch.from_numpy(np.ascontiguousarray(np.tile(sit[order], (4, 1)))).cuda() if name == "input order" else \
                torch.from_numpy(np.ascontiguousarray(np.tile(sit, (4, 1))[np.argsort(np.tile(key, 4), kind="stable")])).cuda()
            mst = timed(lambda: hb.hand_potential_counts(big, "treys", turn_variant=4))
            line += f"   treys turn kernel: {big.shape[0] / mst / 1e3:.1f} M/s"
        print(line, flush=True)
