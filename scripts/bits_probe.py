"""A/B of the flop kernel's undo half: variant 4 (bit-parallel, production) vs 7 (step by step), equality + timing per board texture."""
import json, os, sys
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
sit = hb.random_situations(n, seed=20251019)
d = torch.from_numpy(sit).cuda()
golden = json.load(open(os.path.join(ROOT, "tests", "golden", "gpu_hashes.json")))
r4 = hb.hand_potential_counts(d, "treys", flop_variant=4)
r7 = hb.hand_potential_counts(d, "treys", flop_variant=7)
r5 = hb.hand_potential_counts(d, "treys", flop_variant=5)
bad = (r4 != r7).any(dim=1).nonzero().flatten()
print("golden", golden.get("configs2_first_65536_rows_treys"))
print("variant 4 == 7:", bad.numel() == 0, " 5 == 7:", bool(torch.equal(r5, r7)), " hash4 %016x" % hb.rows_hash(r4), flush=True)
if bad.numel():
    bs = sit[:, 2:5] // 13
    for i in bad[:12].tolist():
        print(i, sit[i].tolist(), "suits", (sit[i, :5] // 13).tolist(), "\n  got ", r4[i].tolist(), "\n  want", r7[i].tolist())
    tex = np.array([max(np.bincount(r, minlength=4)) for r in bs])
    print("bad by texture (max suit count 1/2/3):", np.bincount(tex[bad.cpu().numpy()], minlength=4).tolist(), "of", np.bincount(tex, minlength=4).tolist())
for v in (7, 4, 5, 6):
    ms = timed(lambda: hb.hand_potential_counts(d, "treys", flop_variant=v))
    print(f"mixed variant {v}: {n / ms / 1e3:.2f} M/s", flush=True)
bs = sit[:, 2:5] // 13
tex = np.array([max(np.bincount(r, minlength=4)) for r in bs])
for t, name in ((1, "rainbow"), (2, "two-tone"), (3, "monotone")):
    sel = sit[tex == t]
    sel = np.ascontiguousarray(np.tile(sel, (max(1, n // len(sel)), 1)))
    dd = torch.from_numpy(sel).cuda()
    out = []
    for v in (7, 4):
        ms = timed(lambda: hb.hand_potential_counts(dd, "treys", flop_variant=v), reps=5)
        out.append(f"v{v} {len(sel) / ms / 1e3:.2f} M/s")
    print(f"{name:9s} ({len(sel)} rows): " + "  ".join(out), flush=True)
rows, cyc = hb.flop_phase_cycles(d)
c = cyc.cpu().numpy().astype(np.float64)
print("phase cycle shares [lists+rows, HS+sets, claim+T wait, flush tasks, our-flush entries, descriptors, generic tasks, -]:", (c / c.sum()).round(3).tolist(), " cycles/situation/group:", int(c.sum() / n),
      " rows equal:", bool(torch.equal(rows, r4)))
