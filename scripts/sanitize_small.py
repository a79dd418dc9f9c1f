"""Small run of the production HP kernels for compute-sanitizer (memcheck / racecheck / initcheck / synccheck):
texture cases + a few random flop, turn and river rows in one batch, results checked against the oracle."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dask-holdem-simulator_b200"))
sys.path.insert(0, ROOT)
import holdem_b200 as hb  # noqa: E402
from oracle import oracle as O  # noqa: E402

hb._lib.init(0)
flops = [([14, 1], [31, 32, 33]), ([0, 1], [2, 3, 4]), ([26, 27], [0, 2, 17]), ([12, 25], [38, 51, 0]), ([39, 40], [41, 42, 0]),
         ([0, 1], [5, 18, 31]), ([50, 51], [49, 48, 47])]
turns = [([0, 1], [2, 3, 4, 5]), ([13, 26], [2, 3, 4, 5]), ([12, 11], [2, 3, 4, 18]), ([39, 1], [41, 42, 43, 44]),
         ([0, 13], [2, 3, 15, 16])]
sit = np.concatenate([hb.pack_situations([c[0] for c in flops], [c[1] for c in flops]),
                      hb.pack_situations([c[0] for c in turns], [c[1] for c in turns]),
                      hb.random_situations(20, seed=1), hb.random_situations(20, seed=2, n_board=4),
                      hb.random_situations(8, seed=3, n_board=5)])
want = O.hp_batch(sit, "treys")
got = hb.hand_potential_counts(torch.from_numpy(sit).cuda(), "treys").cpu().numpy().astype(np.uint32)
print("rows:", len(sit), "equal to oracle:", bool(np.array_equal(got, want)))
