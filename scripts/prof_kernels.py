"""Launch one named kernel a few times on its bench-sized workload (for ncu captures):
   python scripts/prof_kernels.py flop4 | flop7 (the step-by-step undo half) | ref5_flop | ref5_turn | ref5_river | hs_ref5 | hs_treys4 | turn4"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dask-holdem-simulator_b200"))
import holdem_b200 as hb  # noqa: E402

which = sys.argv[1]
hb._lib.init(0)
nb = {"flop7": 3, "flop4": 3, "ref5_flop": 3, "ref5_turn": 4, "ref5_river": 5, "hs_ref5": 3, "hs_treys4": 3, "turn4": 4}[which]
n = 1 << 20 if which.startswith("hs_") or which == "turn4" else 65536
d = torch.from_numpy(hb.random_situations(n, seed=20251019, n_board=nb)).cuda()
for _ in range(3):
    if which in ("flop4", "flop7"):
        out = hb.hand_potential_counts(d, "treys", flop_variant=int(which[4]))
    elif which == "turn4":
        out = hb.hand_potential_counts(d, "treys", turn_variant=4)
    elif which.startswith("ref5"):
        out = hb.hand_potential_counts(d, "reference", ref_variant=5)
    elif which == "hs_ref5":
        out = hb.hand_strength_counts(d, "reference")
    else:
        out = hb.hand_strength_counts(d, "treys")
torch.cuda.synchronize()
print("ok", which, n, int(out.sum()))
