"""the UNMODIFIED reference (/root/reference/HandCalculations.py behind stub modules,
oracle/reference_loader.py) timed on this container's CPU cores on situations of the bench workload (configs[2] deals):
HandStrength on 256 flop situations and HandPotential on one situation per core, one process per core.  The results of the
timed calls are compared with the C oracle's literal mode.  Container-only (the reference tree does not travel to the GPU
box); the output is committed as profiles/r1_reference_python_timing_container.json.
"""
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dask-holdem-simulator_b200"))
os.environ["CUDA_VISIBLE_DEVICES"] = ""


def _deals(n):
    import numpy as np
    rng = np.random.default_rng(20251019)
    return np.argsort(rng.random((65536, 52)), axis=1)[:n, :5].astype(int)


def _hs_worker(rows):
    import torch
    torch.set_num_threads(1)
    from oracle.reference_loader import load_reference
    HC = load_reference()
    t0 = time.perf_counter()
    out = [HC.HandStrength(torch.tensor(r[:2]), torch.tensor(r[2:])) for r in rows]
    return time.perf_counter() - t0, out


def _hp_worker(row):
    import torch
    torch.set_num_threads(1)
    from oracle.reference_loader import load_reference
    HC = load_reference()
    t0 = time.perf_counter()
    hp = HC.HandPotential(torch.tensor(row[:2]), torch.tensor(row[2:]))
    return time.perf_counter() - t0, [float(hp[0]), float(hp[1])]


def main():
    from oracle import oracle as O
    cores = len(os.sched_getaffinity(0))
    deals = _deals(256)
    with mp.get_context("spawn").Pool(cores) as pool:
        t0 = time.perf_counter()
        parts = pool.map(_hs_worker, [deals[i::cores].tolist() for i in range(cores)])
        hs_wall = time.perf_counter() - t0
        hs_ok = True
        for i, (dt, vals) in enumerate(parts):
            for r, v in zip(deals[i::cores].tolist(), vals):
                a, t, b = O.hs_counts(r[:2], r[2:], "reference")
                hs_ok = hs_ok and v == O.hs_float(a, t, b)
        t0 = time.perf_counter()
        hp_parts = pool.map(_hp_worker, deals[:cores].tolist())
        hp_wall = time.perf_counter() - t0
    hp_ok = True
    for r, (dt, vals) in zip(deals[:cores].tolist(), hp_parts):
        HP, tot = O.hp_counts(r[:2], r[2:], "reference")
        pp, npot = O.hp_floats(HP, tot)
        hp_ok = hp_ok and vals == [float(pp), float(npot)]
    res = {
        "what": "unmodified reference HandCalculations.py (stub poker/colorama/treys modules), torch CPU, 1 thread per process",
        "cores": cores,
        "HandStrength": {"situations": 256, "wall_s": hs_wall, "situations_per_sec_all_cores": 256 / hs_wall,
                         "cpu_s_per_situation": sum(p[0] for p in parts) / 256, "equal_to_oracle_literal": hs_ok},
        "HandPotential": {"situations": cores, "wall_s": hp_wall, "situations_per_sec_all_cores": cores / hp_wall,
                          "cpu_s_per_situation": sum(p[0] for p in hp_parts) / cores, "equal_to_oracle_literal": hp_ok},
    }
    print(json.dumps(res, indent=1))
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            f.write(json.dumps(res, indent=1) + "\n")


if __name__ == "__main__":
    main()
