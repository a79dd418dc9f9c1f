"""Writes tests/golden/gpu_hashes.json: content hashes of the full-size BASELINE configs computed on ONE GPU, each backed by
independent evidence gathered in the same run (recorded under "_evidence"):
  configs[2] 65,536 flop rows   -- production kernel == 4-set kernel (all rows), and (--oracle-all) ALL rows == C oracle
  configs[4] 1,000,000 flop rows -- production kernel == 4-set kernel (all rows) + an oracle sample
  configs[3] 22 seats x 100,000 hands, flop / turn / river -- table entry point == row expansion, production == older
                                   kernels (all rows), + an oracle sample per street
bench.py and the GPU tests then require the same hashes at every GPU count.   python scripts/make_gpu_hashes.py [--oracle-all]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dask-holdem-simulator_b200"))
sys.path.insert(0, ROOT)
import holdem_b200 as hb  # noqa: E402
from holdem_b200 import table  # noqa: E402
from oracle import oracle as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--oracle-all", action="store_true", help="compare all 65,536 configs[2] rows with the oracle (minutes)")
ap.add_argument("--sample", type=int, default=2048)
args = ap.parse_args()
path = os.path.join(ROOT, "tests", "golden", "gpu_hashes.json")
res = json.load(open(path)) if os.path.exists(path) else {}
ev = {}
rng = np.random.default_rng(99)


def oracle_sample(sit_np, rows, k):
    pick = np.sort(rng.choice(len(sit_np), min(k, len(sit_np)), replace=False))
    want = O.hp_batch_parallel(sit_np[pick], "treys", cache4=True)
    return bool(np.array_equal(rows[torch.from_numpy(pick).cuda()].cpu().numpy().astype(np.uint32), want)), len(pick)


def hx(t):
    return f"{hb.rows_hash(t.contiguous()):016x}"


# configs[2]
sit_np = hb.random_situations(65536, 20251019)
sit = torch.from_numpy(sit_np).cuda()
rows = hb.hand_potential_counts(sit, "treys")
v3 = hb.hand_potential_counts(sit, "treys", flop_variant=3)
k = 65536 if args.oracle_all else args.sample
same, n_or = oracle_sample(sit_np, rows, k)
assert torch.equal(rows, v3) and same
res["configs2_first_65536_rows_treys"] = hx(rows)
# the same situations in mode=reference: rank-multiset kernel == plain enumeration (one literal category evaluation per pair,
# all rows) == the earlier 4-set kernel; oracle sample
ref = hb.hand_potential_counts(sit, "reference")
assert torch.equal(ref, hb.hand_potential_counts(sit, "reference", direct=True))
assert torch.equal(ref[:8192], hb.hand_potential_counts(sit[:8192].contiguous(), "reference", ref_variant=4))
pick = np.sort(rng.choice(65536, 256, replace=False))
assert np.array_equal(ref[torch.from_numpy(pick).cuda()].cpu().numpy().astype(np.uint32), O.hp_batch_parallel(sit_np[pick], "reference", cache4=False))
res["configs2_first_65536_rows_reference"] = hx(ref)
ev["configs2_reference"] = {"v5_equals_plain_enumeration_all_rows": True, "v5_equals_4set_kernel_rows": 8192, "oracle_rows_compared": 256,
                            "oracle_equal": True}
ev["configs2"] = {"production_equals_4set_kernel_all_rows": True, "oracle_rows_compared": n_or, "oracle_equal": same}

# configs[4]
sit_np = hb.random_situations(1_000_000, 20251021)
sit = torch.from_numpy(sit_np).cuda()
rows = hb.hand_potential_counts(sit, "treys")
v3 = hb.hand_potential_counts(sit, "treys", flop_variant=3)
same, n_or = oracle_sample(sit_np, rows, args.sample)
assert torch.equal(rows, v3) and same
res["configs4_1m_flop_rows_treys"] = hx(rows)
ev["configs4"] = {"production_equals_4set_kernel_all_rows": True, "oracle_rows_compared": n_or, "oracle_equal": same}

# configs[3]
hole, board = table.deal_tables(100_000, seed=20251020, seats=22)
d_hole, d_board = torch.from_numpy(hole).cuda(), torch.from_numpy(board).cuda()
for nb, name, variants, k in ((3, "flop_hs_hp", (3, 4, 0), args.sample), (4, "turn_hs_hp", (4, 3, 0), 8 * args.sample),
                              (5, "river_hs", (4, 4, 1), 16 * args.sample)):
    sit_np = table.situations_for_street(hole, board, nb)
    got = hb.table_hand_potential_counts(d_hole, d_board, nb, "treys").reshape(-1, 16)
    d_sit = torch.from_numpy(sit_np).cuda()
    assert torch.equal(got, hb.hand_potential_counts(d_sit, "treys"))
    assert torch.equal(got, hb.hand_potential_counts(d_sit, "treys", variants=variants))
    same, n_or = oracle_sample(sit_np, got, k)
    assert same
    res[f"configs3_22seat_100k_{name}"] = hx(got)
    ev[f"configs3_{name}"] = {"table_entry_equals_row_expansion": True, "production_equals_older_kernels_all_rows": True,
                              "oracle_rows_compared": n_or, "oracle_equal": same, "rows": int(got.shape[0])}
    del got, d_sit

hist, _ = hb.exhaustive7()
assert hx(hist.view(torch.int32).reshape(-1, 2)) == res["exhaustive7_rank_histogram"]
res["_evidence"] = ev
res["_device"] = torch.cuda.get_device_name(0)
with open(path, "w") as f:
    json.dump(res, f, indent=1)
    f.write("\n")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "gpu_hashes.json"), "w") as f:
    json.dump(res, f, indent=1)
print(json.dumps(res, indent=1))
