import os, sys, faulthandler
faulthandler.enable()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dask-holdem-simulator_b200"))
import torch
import holdem_b200 as hb
sit = torch.from_numpy(hb.random_situations(64, seed=1)).cuda()
print("hp ok", torch.ops.holdem.hp(sit, 0).shape, flush=True)
for name, fn in (("cpu tensor", lambda: torch.ops.holdem.hp(sit.cpu(), 0)),
                 ("bad shape", lambda: torch.ops.holdem.hp(sit[:, :7].contiguous(), 0)),
                 ("bad mode", lambda: torch.ops.holdem.hp(sit, 7)),
                 ("bad n_board", lambda: torch.ops.holdem.table_hp(sit[:, :2].reshape(4, 16, 2).contiguous(), sit[:4, :5].contiguous(), 2, 0))):
    print("trying", name, flush=True)
    try:
        fn()
        print("  no exception", flush=True)
    except Exception as e:
        print("  raised", type(e).__name__, str(e)[:150].replace("\n", " | "), flush=True)
print("done", flush=True)
