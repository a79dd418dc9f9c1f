"""Reference-mode kernel A/B: variant 5 (production) vs 4 (4-set walk) on a sample, content hash of the 65,536 bench rows, timing per street."""
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

golden = json.load(open(os.path.join(ROOT, "tests", "golden", "gpu_hashes.json")))
n = 65536
d = torch.from_numpy(hb.random_situations(n, seed=20251019)).cuda()
r5 = hb.hand_potential_counts(d, "reference")
print("hash %016x" % hb.rows_hash(r5), "committed", golden.get("configs2_first_65536_rows_reference"))
small = d[:8192]
print("v5 == 4-set kernel on 8192 rows:", bool(torch.equal(hb.hand_potential_counts(small, "reference", ref_variant=5), hb.hand_potential_counts(small, "reference", ref_variant=4))))
print("v5 == plain enumeration on 2048 rows:", bool(torch.equal(r5[:2048], hb.hand_potential_counts(d[:2048], "reference", direct=True))))
for nb in (3, 4, 5):
    dd = torch.from_numpy(hb.random_situations(n, seed=20251019 + nb, n_board=nb)).cuda()
    ms = timed(lambda: hb.hand_potential_counts(dd, "reference"))
    x = hb.hand_potential_counts(dd[:4096], "reference")
    print(f"n_board {nb}: {n / ms / 1e3:.2f} M/s  == plain enumeration (4096 rows): {bool(torch.equal(x, hb.hand_potential_counts(dd[:4096], 'reference', direct=True)))}", flush=True)
