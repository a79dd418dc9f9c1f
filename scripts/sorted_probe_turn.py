"""Host-sorted work order probes for the Treys-mode turn kernel."""
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

n = 1 << 20
sit = hb.random_situations(n, seed=4, n_board=4)
bs = sit[:, 2:6] // 13
hs = sit[:, 0:6] // 13
cnt_b = np.stack([(bs == s).sum(axis=1) for s in range(4)], axis=1)
cnt_h = np.stack([(hs == s).sum(axis=1) for s in range(4)], axis=1)
bmax, hmax = cnt_b.max(axis=1), cnt_h.max(axis=1)
nsuits = (cnt_b > 0).sum(axis=1)
two_two = ((cnt_b == 2).sum(axis=1) == 2).astype(int)
keys = {
    "input order": np.arange(n),
    "board max suit desc": -bmax,
    "bmax desc, our max desc": -bmax * 10 - hmax,
    "bmax desc, two-two, our max": -bmax * 100 - two_two * 10 - hmax,
    "suit pattern (sorted counts)": np.array([hash(tuple(sorted(r))) % 1000 for r in cnt_b.tolist()]),
}
out = torch.empty((n, 16), dtype=torch.int32, device="cuda")
for name, key in keys.items():
    order = np.argsort(key, kind="stable")
    d = torch.from_numpy(np.ascontiguousarray(sit[order])).cuda()
    ms = timed(lambda: hb.hand_potential_counts(d, "treys", out=out, turn_variant=6))
    print(f"{name:32s}: {ms:.4f} ms = {n / ms / 1e3:.1f} M situations/s", flush=True)

d = torch.from_numpy(sit).cuda()
for v, name in ((6, "library, input order (variant 6)"), (4, "library, own ordering pass")):
    ms = timed(lambda: hb.hand_potential_counts(d, "treys", out=out, turn_variant=v))
    print(f"{name:32s}: {ms:.4f} ms = {n / ms / 1e3:.1f} M situations/s", flush=True)
ms = timed(lambda: hb.hand_potential_counts(d, "treys", out=out))
print(f"{'holdem_hp_batch (all kernels)':32s}: {ms:.4f} ms = {n / ms / 1e3:.1f} M situations/s", flush=True)
ref = hb.hand_potential_counts(d, "treys", turn_variant=3)
print("equal to the 3-set kernel:", bool(torch.equal(ref, hb.hand_potential_counts(d, "treys"))), bool(torch.equal(ref, hb.hand_potential_counts(d, "treys", turn_variant=4))), bool(torch.equal(ref, hb.hand_potential_counts(d, "treys", turn_variant=6))))
