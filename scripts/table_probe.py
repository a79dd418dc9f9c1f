"""Timing probe: whole-table entry point vs the same rows through holdem_hp_batch (configs[3] data)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dask-holdem-simulator_b200"))
import holdem_b200 as hb
from holdem_b200 import table

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

hands = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
hole, board = table.deal_tables(hands, seed=20251020, seats=22)
d_hole, d_board = torch.from_numpy(hole).cuda(), torch.from_numpy(board).cuda()
for nb in (3, 4, 5):
    rows = torch.from_numpy(table.situations_for_street(hole, board, nb)).cuda()
    out = torch.empty((rows.shape[0], 16), dtype=torch.int32, device="cuda")
    t_rows = timed(lambda: hb.hand_potential_counts(rows, "treys", out=out))
    t_tab = timed(lambda: hb.table_hand_potential_counts(d_hole, d_board, nb, "treys"))
    perm = torch.randperm(rows.shape[0], device="cuda")
    shuffled = rows[perm].contiguous()
    t_shuf = timed(lambda: hb.hand_potential_counts(shuffled, "treys", out=out))
    print(f"n_board {nb}: rows {t_rows:.2f} ms, table entry {t_tab:.2f} ms, same rows shuffled {t_shuf:.2f} ms", flush=True)
rnd = torch.from_numpy(hb.random_situations(2200000, seed=5)).cuda()
out = torch.empty((rnd.shape[0], 16), dtype=torch.int32, device="cuda")
print("2.2M independent random flop rows:", timed(lambda: hb.hand_potential_counts(rnd, "treys", out=out)), "ms")
