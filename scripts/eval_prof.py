import sys, os
sys.path.insert(0, "/root/repo/dask-holdem-simulator_b200"); sys.path.insert(0, "/root/repo")
import torch, holdem_b200 as hb
n = 1 << 25
hands = torch.empty((n, 8), dtype=torch.uint8, device="cuda"); hands.fill_(255)
for s in range(0, n, 1 << 21):
    hands[s:s + (1 << 21), :7] = torch.rand((1 << 21, 52), device="cuda").topk(7, dim=1).indices.to(torch.uint8)
out = torch.empty(n, dtype=torch.int16, device="cuda")
for _ in range(3): hb.evaluate(hands, 7, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): hb.evaluate(hands, 7, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("eval7", n / ms / 1e6, "G evals/s", ms, "ms", 10 * n / ms / 1e6, "GB/s")
