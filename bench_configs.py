#!/usr/bin/env python
"""bench_configs.py -- the BASELINE.json configs that are parity-test cases rather than the bench.py line, timed on
the GPU and written as one JSON document (kept under profiles/).  bench.py stays the driver-facing contract.

  python bench_configs.py [--hands 100000] [--ref-situations 4096] [--out profiles/xyz.json]
  torchrun ... bench_configs.py --config5 [--situations 1000000]       (1M flop situations sharded, strong scaling)
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "dask-holdem-simulator_b200")
for _p in (PKG, ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np
import torch

import holdem_b200 as hb


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def table_situations(n_hands, seed=20251020, seats=22):
    """Config 4: one shuffle per hand, seat i holds cards 2i,2i+1, board = cards 44..48."""
    rng = np.random.default_rng(seed)
    out = {}
    chunks = {3: [], 4: [], 5: []}
    step = 20000
    for s in range(0, n_hands, step):
        m = min(step, n_hands - s)
        deck = np.argsort(rng.random((m, 52)), axis=1).astype(np.uint8)
        hole = deck[:, :2 * seats].reshape(m, seats, 2)
        for nb in (3, 4, 5):
            rows = np.full((m, seats, 8), 0xFF, dtype=np.uint8)
            rows[:, :, 0:2] = hole
            rows[:, :, 2:2 + nb] = deck[:, None, 44:44 + nb]
            chunks[nb].append(rows.reshape(m * seats, 8))
    for nb in (3, 4, 5):
        out[nb] = np.concatenate(chunks[nb])
    return out


def config5(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG") and not os.environ.get("NCCL_DEBUG_FILE"):
            os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"      # keep stdout to the JSON line, the log stays visible
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.situations
    sit = torch.from_numpy(hb.random_situations(n, seed=20251021)).cuda()
    for _ in range(2):
        full = hb.hand_potential_counts_sharded(sit, "treys")
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        full = hb.hand_potential_counts_sharded(sit, "treys")
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    r = full.cpu().numpy().astype(np.int64)
    ok = bool((r[:, 12:15].sum(axis=1) == 1081).all()
              and (r[:, 3:12].reshape(-1, 3, 3).sum(axis=2) == 990 * r[:, 12:15]).all())
    content_hash = f"{hb.rows_hash(full.contiguous()):016x}"
    if rank == 0:
        print(json.dumps({"config": "configs[4]: 1M random flop situations HS+HP sharded over the GPUs, every rank receives all rows (kernels store into all ranks' symmetric-memory buffers over NVLink; NCCL all-gather as fallback)",
                          "n_gpus": world, "situations": n, "ms": float(ms.item()),
                          "situations_per_sec": n / (float(ms.item()) * 1e-3), "scaling": "strong",
                          "properties_ok": ok, "content_hash": content_hash}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hands", type=int, default=100000)
    ap.add_argument("--ref-situations", type=int, default=4096)
    ap.add_argument("--situations", type=int, default=1000000)
    ap.add_argument("--config5", action="store_true")
    ap.add_argument("--out", default=None)
    ap.add_argument("--oracle-fraction", type=float, default=0.0,
                    help="config 4: fraction of the rows of every street compared with the C oracle (0.01 = 1 %%, ~2 min of CPU)")
    ap.add_argument("--python-baseline", action="store_true", help="also time the reference-shaped pure-Python path (config 1)")
    ap.add_argument("--python-baseline-opponents", type=int, default=40,
                    help="opponent holdings of the 1081 to time (each = 990 runouts x 2 evaluate calls)")
    args = ap.parse_args()
    if args.config5:
        return config5(args)

    torch.cuda.set_device(0)
    hb._lib.init(0)
    res = {"device": torch.cuda.get_device_name(0)}

    # ---- config 1: one flop situation through the drop-in module (latency, includes H2D/D2H and Python) ------------------
    import HandCalculations as HC
    our_hs, board_hs = torch.tensor([0, 13]), torch.tensor([26, 27, 28, 29, 30])
    our_hp, board_hp = torch.tensor([14, 1]), torch.tensor([31, 32, 33])
    c1 = {}
    for mode in ("treys", "reference"):
        HC.set_mode(mode)
        HC.HandPotential(our_hp, board_hp)
        t0 = time.perf_counter()
        for _ in range(20):
            hs = HC.HandStrength(our_hs, board_hs)
        t_hs = (time.perf_counter() - t0) / 20
        t0 = time.perf_counter()
        for _ in range(20):
            hp = HC.HandPotential(our_hp, board_hp)
            float(hp[0])
        t_hp = (time.perf_counter() - t0) / 20
        c1[mode] = {"HandStrength": hs, "HandPotential": [float(x) for x in hp],
                    "HandStrength_ms": 1e3 * t_hs, "HandPotential_ms": 1e3 * t_hp}
    HC.set_mode("treys")
    res["config1_single_flop_dropin"] = c1
    if args.python_baseline:
        # the same situation through the reference-shaped Python path on one host core: 1081 x 990 runouts, two pure-Python
        # Treys evaluate() calls each (oracle/py_treys.py restates treys.Evaluator; treys itself cannot be installed offline)
        import itertools
        from oracle.py_treys import PyEvaluator, card_from_path
        ev = PyEvaluator()
        our, board = [14, 1], [31, 32, 33]
        unseen = [c for c in range(52) if c not in our + board]
        tc = {c: card_from_path(c) for c in range(52)}
        t0 = time.perf_counter()
        ours_now = ev.evaluate([tc[c] for c in board], [tc[c] for c in our])
        hp = [[0, 0, 0], [0, 0, 0], [0, 0, 0]]
        n_opp = 0
        budget = args.python_baseline_opponents
        for p in itertools.combinations(unseen, 2):
            if n_opp == budget:
                break
            n_opp += 1
            opp_now = ev.evaluate([tc[c] for c in board], [tc[c] for c in p])
            idx = 0 if ours_now < opp_now else (1 if ours_now == opp_now else 2)
            rest = [c for c in unseen if c not in p]
            for q in itertools.combinations(rest, 2):
                b5 = [tc[c] for c in board + list(q)]
                a, b = ev.evaluate(b5, [tc[c] for c in our]), ev.evaluate(b5, [tc[c] for c in p])
                hp[idx][0 if a < b else (1 if a == b else 2)] += 1
        dt = time.perf_counter() - t0
        res["config1_python_path_one_core"] = {
            "opponent_holdings_timed": n_opp, "seconds": dt, "seconds_per_situation_extrapolated": dt * 1081 / n_opp,
            "evaluate_calls": n_opp * (1 + 2 * 990) + 1,
            "what": "HandPotential loop shape of the reference with a pure-Python Treys evaluator, testing/ExampleHS.py situation"}

    # ---- config 2: exhaustive sweep -----------------------------------------------------------------------------------------
    ms = timed(lambda: hb.exhaustive7(0), reps=5)
    hist, sums = hb.exhaustive7(0)
    res["config2_exhaustive7"] = {"hands": 133784560, "ms": ms, "evals_per_sec": 133784560 / (ms * 1e-3),
                                  "sum_rank": int(sums[0]), "sum_rank_sq": int(sums[1]),
                                  "distinct_ranks": int((hist > 0).sum()),
                                  "ok": sums.cpu().tolist() == [547965983972, 2665153084455708]}

    # ---- per-kernel rates on random situations ----------------------------------------------------------------------------------
    rates = {}
    n = args.ref_situations
    for nb in (3, 4, 5):
        sit = torch.from_numpy(hb.random_situations(max(n, 65536), seed=31 + nb, n_board=nb)).cuda()
        for mode in ("treys", "reference"):
            ms = timed(lambda: hb.hand_strength_counts(sit, mode))
            rates[f"hs_{mode}_board{nb}"] = {"situations": sit.shape[0], "ms": ms,
                                             "situations_per_sec": sit.shape[0] / (ms * 1e-3)}
        small = sit[:n].contiguous()
        for mode in ("treys", "reference"):
            ms = timed(lambda: hb.hand_potential_counts(small, mode), reps=2)
            rates[f"hp_{mode}_board{nb}"] = {"situations": n, "ms": ms, "situations_per_sec": n / (ms * 1e-3)}
        if nb == 3:
            ms = timed(lambda: hb.hand_potential_counts(small, "treys", direct=True), reps=2)
            rates["hp_treys_board3_plain_enumeration"] = {"situations": n, "ms": ms,
                                                          "situations_per_sec": n / (ms * 1e-3)}
    for nc in (5, 6, 7):
        hands = torch.from_numpy(hb.random_situations(1 << 22, seed=nc, n_board=nc - 2)).cuda()
        out = torch.empty(hands.shape[0], dtype=torch.int16, device="cuda")
        ms = timed(lambda: hb.evaluate(hands, nc, out=out), reps=10)
        rates[f"eval{nc}"] = {"hands": hands.shape[0], "ms": ms, "evals_per_sec": hands.shape[0] / (ms * 1e-3)}
    cards = torch.randint(0, 52, (1 << 22, 16), dtype=torch.uint8, device="cuda")
    for nc in (7, 9):
        ms = timed(lambda: hb.category(cards, nc), reps=10)
        rates[f"category{nc}"] = {"hands": cards.shape[0], "ms": ms, "hands_per_sec": cards.shape[0] / (ms * 1e-3)}
    res["kernel_rates"] = rates

    # ---- config 4: 22-seat table, every seat on flop / turn / river, through the whole-table entry point -----------------------------
    from holdem_b200 import table
    hole, board = table.deal_tables(args.hands, seed=20251020, seats=22)
    d_hole, d_board = torch.from_numpy(hole).cuda(), torch.from_numpy(board).cuda()
    golden = {}
    try:
        golden = json.load(open(os.path.join(ROOT, "tests", "golden", "gpu_hashes.json")))
    except Exception:
        pass
    c4 = {"hands": args.hands, "seats": 22, "situations_per_street": args.hands * 22, "oracle_fraction": args.oracle_fraction}
    total_ms = 0.0
    rng = np.random.default_rng(4)
    for nb, name in ((3, "flop_hs_hp"), (4, "turn_hs_hp"), (5, "river_hs")):
        fn = lambda: hb.table_hand_potential_counts(d_hole, d_board, nb, "treys")
        ms = timed(fn, reps=2, warm=1)
        total_ms += ms
        out = fn().reshape(-1, 16)
        r = out.cpu().numpy().astype(np.int64)
        m = 52 - 2 - nb
        runouts = {3: 990, 4: 44, 5: 0}[nb]
        ok = bool((r[:, 12:15].sum(axis=1) == m * (m - 1) // 2).all()
                  and (r[:, 3:12].reshape(-1, 3, 3).sum(axis=2) == runouts * r[:, 12:15]).all())
        entry = {"ms": ms, "situations_per_sec": out.shape[0] / (ms * 1e-3), "properties_ok": ok,
                 "content_hash": f"{hb.rows_hash(out):016x}"}
        key = f"configs3_22seat_100k_{name}"
        if args.hands == 100000 and key in golden:
            entry["hash_equals_committed"] = entry["content_hash"] == golden[key]
        if args.oracle_fraction > 0:
            from oracle import oracle as O
            n_or = max(1, int(out.shape[0] * args.oracle_fraction))
            pick = np.sort(rng.choice(out.shape[0], n_or, replace=False))
            t0 = time.perf_counter()
            want = O.hp_batch_parallel(table.situations_for_street(hole, board, nb)[pick], "treys", cache4=True)
            entry["oracle_rows_compared"] = n_or
            entry["oracle_seconds"] = time.perf_counter() - t0
            entry["oracle_equal"] = bool(np.array_equal(r[pick].astype(np.uint32), want))
        c4[name] = entry
    c4["total_ms"] = total_ms
    c4["hands_per_sec"] = args.hands / (total_ms * 1e-3)
    # showdown of every table (GameBoard.determine_winner): one fused kernel vs the evaluator launch + torch reductions
    ms_f = timed(lambda: hb.table_showdown(d_hole, d_board), reps=5, warm=2)
    ms_e = timed(lambda: (lambda r: (table.winners(r, "best"), table.winners(r, "reference")))(table.showdown_ranks(d_hole, d_board)),
                 reps=5, warm=2)
    ranks_f, best_f, ref_f = table.showdown(d_hole, d_board)
    ranks_e = table.showdown_ranks(d_hole, d_board)
    c4["showdown"] = {"fused_kernel_ms": ms_f, "evaluator_plus_torch_ms": ms_e, "hands_per_sec": args.hands / (ms_f * 1e-3),
                      "equal": bool(torch.equal(ranks_f, ranks_e) and torch.equal(best_f, table.winners(ranks_e, "best"))
                                    and torch.equal(ref_f, table.winners(ranks_e, "reference"))),
                      "split_pots": int((best_f.sum(dim=1) > 1).sum())}
    res["config4_22_seat_table"] = c4

    text = json.dumps(res, indent=1)
    print(text)
    if args.out:
        with open(args.out, "w") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main()
