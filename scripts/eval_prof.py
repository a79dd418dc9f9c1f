"""Times the large-batch evaluator (5/6/7 cards, 2^25 hands) and checks it row for row against the strict kernel,
including rows with out-of-range bytes.  GPU only."""
import sys
sys.path.insert(0, "/root/repo/dask-holdem-simulator_b200"); sys.path.insert(0, "/root/repo")
import torch, holdem_b200 as hb
n = 1 << 25
for nc in ([int(a) for a in sys.argv[1:]] or (7, 6, 5)):
    hands = torch.empty((n, 8), dtype=torch.uint8, device="cuda"); hands.fill_(255)
    for s in range(0, n, 1 << 21):
        hands[s:s + (1 << 21), :nc] = torch.rand((1 << 21, 52), device="cuda").topk(nc, dim=1).indices.to(torch.uint8)
    bad = torch.randint(0, n, (4096,), device="cuda")
    hands[bad, torch.randint(0, nc, (4096,), device="cuda")] = torch.randint(52, 256, (4096,), device="cuda").to(torch.uint8)
    out = torch.empty(n, dtype=torch.int16, device="cuda")
    for _ in range(3): hb.evaluate(hands, nc, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): hb.evaluate(hands, nc, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    ref = hb.evaluate(hands, nc, strict=True)
    print(f"eval{nc}", n / ms / 1e6, "G evals/s", ms, "ms", 10 * n / ms / 1e6, "GB/s", "equal_strict", bool(torch.equal(out, ref)),
          "zeros", int((out == 0).sum()), flush=True)
    del hands, out, ref
