"""Launch the production flop kernel a few times on the bench workload (for ncu captures)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dask-holdem-simulator_b200"))
import holdem_b200 as hb  # noqa: E402

hb._lib.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
d = torch.from_numpy(hb.random_situations(n, seed=20251019)).cuda()
out = torch.empty((n, 16), dtype=torch.int32, device="cuda")
for _ in range(3):
    hb.hand_potential_counts(d, "treys", out=out, flop_variant=4)
torch.cuda.synchronize()
print("ok", int(out[:, 12:15].sum()))
