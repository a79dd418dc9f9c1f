import os, sys, time
import numpy as np, torch
sys.path.insert(0, '/root/repo/dask-holdem-simulator_b200')
import holdem_b200 as hb
hb._lib.init(0)
n = 65536
sit = hb.random_situations(n, seed=20251019)
sit[5, 1] = sit[5, 0]
pin_in = torch.empty((n, 8), dtype=torch.uint8, pin_memory=True); pin_in.numpy()[:] = sit
pin_out = torch.zeros((n, 16), dtype=torch.int32, pin_memory=True)
out = pin_out.numpy().view(np.uint32)
try:
    hb.hand_potential_counts_host(pin_in.numpy(), "treys", out=out)
except hb.HoldemError as e:
    print("bad row reported:", e.code)
pin_in.numpy()[5, 1] = (sit[5, 0] + 1) % 52 if (sit[5, 0] + 1) % 52 not in sit[5, 2:5] else (sit[5,0] + 7) % 52
ref = hb.hand_potential_counts(torch.from_numpy(pin_in.numpy().copy()).cuda(), "treys").cpu().numpy().astype(np.uint32)
for name in ("zero-copy",):
    hb.hand_potential_counts_host(pin_in.numpy(), "treys", out=out)
    print(name, "equal to device entry:", np.array_equal(out, ref))
    t0 = time.perf_counter()
    for _ in range(50):
        hb.hand_potential_counts_host(pin_in.numpy(), "treys", out=out)
    dt = (time.perf_counter() - t0) / 50
    print(f"{name}: {dt * 1e3:.3f} ms per call = {n / dt / 1e6:.2f} M situations/s")
for mode in ("reference",):
    sub = pin_in.numpy()[:2048]
    o2 = pin_out.numpy().view(np.uint32)[:2048]
    hb.hand_potential_counts_host(sub, mode, out=o2)
    r2 = hb.hand_potential_counts(torch.from_numpy(sub.copy()).cuda(), mode).cpu().numpy().astype(np.uint32)
    print(mode, "equal:", np.array_equal(o2, r2))
