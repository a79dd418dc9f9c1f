"""Times the exhaustive 7-card sweep (BASELINE.json configs[1]) and checks its checksums.  HOLDEM_EXHAUSTIVE_KERNEL=1
selects the older chunk-per-thread kernel.  GPU only."""
import os, sys
sys.path.insert(0, "/root/repo/dask-holdem-simulator_b200"); sys.path.insert(0, "/root/repo")
import torch, holdem_b200 as hb
hist, sums = hb.exhaustive7()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    hist, sums = hb.exhaustive7()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
h = hist.cpu().numpy()
print("kernel", os.environ.get("HOLDEM_EXHAUSTIVE_KERNEL", "tuple"), "ms", ms, "G evals/s", 133784560 / ms / 1e6,
      "sums_ok", sums.cpu().tolist() == [547965983972, 2665153084455708], "hands", int(h.sum()), "distinct", int((h > 0).sum()),
      "royal", int(h[1]))
