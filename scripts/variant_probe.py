"""A/B of differently configured builds of the library (files under _variants/): each takes the production file's place for one
process (run on the GPU box's scratch copy of the repo)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys
sys.path.insert(0, os.path.join(%r, "dask-holdem-simulator_b200"))
import torch
import holdem_b200 as hb
def timed(fn, reps=30):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
n = 65536
d = torch.from_numpy(hb.random_situations(n, seed=20251019)).cuda()
out = torch.empty((n, 16), dtype=torch.int32, device="cuda")
ms = min(timed(lambda: hb.hand_potential_counts(d, "treys", out=out, flop_variant=4)) for _ in range(3))
print(os.path.basename(sys.argv[1]), f"{ms:.4f} ms = {n / ms / 1e3:.2f} M situations/s  hash {hb.rows_hash(out):016x}", flush=True)
''' % ROOT
libs = [os.path.join(ROOT, "dask-holdem-simulator_b200", "holdem_b200", "libholdem_b200.so")]
vd = os.path.join(ROOT, "_variants")
libs += sorted(os.path.join(vd, f) for f in os.listdir(vd) if f.endswith(".so")) + libs[:1]
import shutil
prod = libs[0]
backup = prod + ".orig"
shutil.copy(prod, backup)
try:
    for lib in libs:
        if lib != prod:
            shutil.copy(lib, prod)          # the variant takes the production file's place for this process
        else:
            shutil.copy(backup, prod)
        subprocess.run([sys.executable, "-c", code, lib])
finally:
    shutil.copy(backup, prod)
    os.remove(backup)
