"""Randomised parity soak on one GPU: the production kernels against the independent implementations on far more rows than the
test suite uses (different seeds), every street, both modes.  Writes one JSON document (kept under profiles/)."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dask-holdem-simulator_b200"))
import holdem_b200 as hb

res = {"device": torch.cuda.get_device_name(0), "checks": []}

def check(name, a, b, rows):
    ok = bool(torch.equal(a, b))
    res["checks"].append({"what": name, "rows": rows, "equal": ok})
    print(name, rows, ok, flush=True)
    return ok

t0 = time.time()
ok = True
for seed in (101, 102, 103, 104):
    n = 1 << 20
    d = torch.from_numpy(hb.random_situations(n, seed=seed)).cuda()
    prod = hb.hand_potential_counts(d, "treys")
    ok &= check(f"flop treys: production == 4-set kernel (seed {seed})", prod, hb.hand_potential_counts(d, "treys", flop_variant=3), n)
    ok &= check(f"flop treys: production == no closed form / input order (seed {seed})", hb.hand_potential_counts(d, "treys", flop_variant=5),
                hb.hand_potential_counts(d, "treys", flop_variant=6), n)
    ok &= check(f"flop treys: HS kernel == HP by-product (seed {seed})", hb.hand_strength_counts(d, "treys"), prod[:, :3].contiguous(), n)
    d4 = torch.from_numpy(hb.random_situations(n, seed=seed + 50, n_board=4)).cuda()
    prod4 = hb.hand_potential_counts(d4, "treys")
    ok &= check(f"turn treys: production == 3-set kernel (seed {seed})", prod4, hb.hand_potential_counts(d4, "treys", turn_variant=3), n)
    ok &= check(f"turn treys: HS kernel == HP by-product (seed {seed})", hb.hand_strength_counts(d4, "treys"), prod4[:, :3].contiguous(), n)
for nb in (3, 4, 5):
    n = 1 << 19
    d = torch.from_numpy(hb.random_situations(n, seed=200 + nb, n_board=nb)).cuda()
    v5 = hb.hand_potential_counts(d, "reference")
    ok &= check(f"reference, {nb}-card boards: rank-multiset kernel == plain enumeration", v5, hb.hand_potential_counts(d, "reference", direct=True), n)
    ok &= check(f"reference, {nb}-card boards: rank-pair HS == HP by-product == per-holding HS", hb.hand_strength_counts(d, "reference"),
                hb.hand_strength_counts(d, "reference", variant=1), n)
    ok &= check(f"reference, {nb}-card boards: HS == HP rows", hb.hand_strength_counts(d, "reference"), v5[:, :3].contiguous(), n)
n = 1 << 19
d = torch.from_numpy(hb.random_situations(n, seed=300, n_board=5)).cuda()
ok &= check("river treys: rank-pair HS rows == plain enumeration", hb.hand_potential_counts(d, "treys"), hb.hand_potential_counts(d, "treys", direct=True), n)
mixed = np.concatenate([hb.random_situations(300000, seed=400 + nb, n_board=nb) for nb in (3, 4, 5)])
mixed = mixed[np.random.default_rng(5).permutation(len(mixed))]
dm = torch.from_numpy(mixed).cuda()
ok &= check("mixed streets treys: production == (3, 3, 1) older kernels", hb.hand_potential_counts(dm, "treys"),
            hb.hand_potential_counts(dm, "treys", variants=(3, 3, 1)), len(mixed))
res["all_equal"] = bool(ok)
res["seconds"] = time.time() - t0
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "r2_parity_soak.json"), "w"), indent=1)
print("ALL EQUAL" if ok else "MISMATCH", res["seconds"])
sys.exit(0 if ok else 1)
