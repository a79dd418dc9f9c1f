"""A/B of the flop kernels on the GPU: correctness of every variant against the oracle on a small sample, then CUDA-event
timings on the 65,536-situation bench workload (and per board texture)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dask-holdem-simulator_b200"))
sys.path.insert(0, ROOT)
import holdem_b200 as hb  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    hb._lib.init(0)
    variants = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "4,3").split(",")]
    from oracle import oracle as O
    sit = hb.random_situations(64, seed=5)
    want = O.hp_batch(sit, "treys")
    d = torch.from_numpy(sit).cuda()
    for v in variants:
        got = hb.hand_potential_counts(d, "treys", flop_variant=v).cpu().numpy().astype(np.uint32)
        bad = np.nonzero((got != want).any(axis=1))[0]
        print(f"variant {v}: {len(bad)} mismatching rows of {len(sit)}", flush=True)
        for b in bad[:3]:
            print("  sit", sit[b].tolist(), "\n  got ", got[b].tolist(), "\n  want", want[b].tolist())
    n = 65536
    sit = hb.random_situations(n, seed=20251019)
    d = torch.from_numpy(sit).cuda()
    suits = sit[:, 2:5] // 13
    nsuit = np.array([len(set(r)) for r in suits.tolist()])
    ref = None
    for v in variants:
        out = torch.empty((n, 16), dtype=torch.int32, device="cuda")
        ms = timed(lambda: hb.hand_potential_counts(d, "treys", out=out, flop_variant=v))
        print(f"variant {v}: {ms:.3f} ms per 65,536 situations = {n / ms / 1e3:.3f} M situations/s", flush=True)
        if ref is None:
            ref = out.clone()
        else:
            print("   equal to first variant:", bool(torch.equal(ref, out)))
    _, cyc = hb.flop_phase_cycles(d)
    cyc = cyc.cpu().numpy().astype(np.float64)
    print("v4 phase cycles per situation (tables, W+T, descriptors, tasks):", (cyc / n).round(0).tolist(),
          "share:", (cyc / cyc.sum()).round(3).tolist(), flush=True)
    for name, k in (("monotone", 1), ("two-tone", 2), ("rainbow", 3)):
        sub = np.ascontiguousarray(sit[nsuit == k])
        reps = max(1, 16384 // len(sub))
        sub = np.concatenate([sub] * reps)[:16384] if len(sub) < 16384 else sub[:16384]
        ds = torch.from_numpy(np.ascontiguousarray(sub)).cuda()
        _, cyc = hb.flop_phase_cycles(ds)
        print(f"  {name:9s} phase cycles per situation:", (cyc.cpu().numpy() / len(sub)).round(0).tolist())
        for v in variants:
            ms = timed(lambda: hb.hand_potential_counts(ds, "treys", flop_variant=v))
            print(f"  {name:9s} ({(nsuit == k).mean() * 100:.1f}% of deals) variant {v}: {len(sub) / ms / 1e3:.3f} M situations/s")


if __name__ == "__main__":
    main()
