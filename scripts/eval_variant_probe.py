"""Upper-bound experiments on the batched evaluator: differently built libraries (files under _variants/) take the production
file's place for one process each; 7-card evaluator timing on 2^25 hands."""
import os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys
sys.path.insert(0, os.path.join(%r, "dask-holdem-simulator_b200"))
import torch
import holdem_b200 as hb
n = 1 << 25
hands = torch.empty((n, 8), dtype=torch.uint8, device="cuda"); hands.fill_(255)
g = torch.Generator(device="cuda"); g.manual_seed(1)
for s in range(0, n, 1 << 21):
    hands[s:s + (1 << 21), :7] = torch.rand((1 << 21, 52), device="cuda", generator=g).topk(7, dim=1).indices.to(torch.uint8)
out = torch.empty(n, dtype=torch.int16, device="cuda")
for _ in range(3): hb.evaluate(hands, 7, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): hb.evaluate(hands, 7, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(sys.argv[1], f"{n / ms / 1e6:.1f} G evals/s", flush=True)
''' % ROOT
prod = os.path.join(ROOT, "dask-holdem-simulator_b200", "holdem_b200", "libholdem_b200.so")
vd = os.path.join(ROOT, "_variants")
libs = [prod] + sorted(os.path.join(vd, f) for f in os.listdir(vd) if f.endswith(".so")) + [prod]
backup = prod + ".orig"
shutil.copy(prod, backup)
try:
    for lib in libs:
        shutil.copy(backup if lib == prod else lib, prod)
        subprocess.run([sys.executable, "-c", code, os.path.basename(lib)])
finally:
    shutil.copy(backup, prod)
    os.remove(backup)
