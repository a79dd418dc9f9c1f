"""CPU-only checks of the product's host side: the C-ABI library loads and exports every symbol the header declares,
the generated tables pass their self-check, entry points fail loudly without a GPU, converters, sharding arithmetic,
and the drop-in module's surface.  No compute call is made without a GPU."""
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT

import holdem_b200 as hb
from holdem_b200 import _lib


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "holdem_b200.h")).read()
    declared = set(re.findall(r"\b(holdem_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.holdem_abi_version() == _lib.ABI_VERSION == 2


def test_tables_selfcheck():
    assert _lib.load().holdem_selftest_tables() == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    lib = _lib.load()
    assert lib.holdem_init(0) == _lib.ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.holdem_last_error()
    with pytest.raises(_lib.HoldemError):
        _lib.init(0)
    with pytest.raises(RuntimeError):
        hb.hand_potential_counts(torch.zeros((1, 8), dtype=torch.uint8), "treys")
    import HandCalculations as HC
    with pytest.raises(RuntimeError):
        HC.HandStrength(torch.tensor([0, 1]), torch.tensor([2, 3, 4]))
    with pytest.raises(RuntimeError):
        HC.CalculateHandRankValue([0, 1], [10, 11, 12, 13, 14])
    # the whole-table entry points refuse host tensors too (no CPU path behind them)
    hole, board = torch.zeros((1, 22, 2), dtype=torch.uint8), torch.zeros((1, 5), dtype=torch.uint8)
    with pytest.raises(RuntimeError):
        hb.table_showdown(hole, board)
    with pytest.raises(RuntimeError):
        hb.table_hand_potential_counts(hole, board, 3, "treys")


def test_encoding_round_trips():
    for c in range(52):
        s = hb.str_from_path(c)
        assert hb.path_from_str(s) == c
        assert hb.path_from_treys(hb.treys_from_path(c)) == c
        assert hb.path_from_rank_major(hb.rank_major_from_path(c)) == c
    assert hb.treys_from_str("Kd") == 0x08004B25 and hb.treys_from_str("5s") == 0x00081307
    assert hb.treys_from_str("Jc") == 0x0200891D
    assert hb.path_from_str("2c") == 0 and hb.path_from_str("3d") == 14 and hb.path_from_str("As") == 51
    with pytest.raises(ValueError):
        hb.path_from_treys(12345)
    # utils.card_to_int: rank * 4 + suit
    assert hb.rank_major_from_path(hb.path_from_str("Ah")) == 12 * 4 + 2


def test_pack_and_random_situations():
    rows = hb.pack_situations([[14, 1]], [[31, 32, 33]])
    assert rows.tolist() == [[14, 1, 31, 32, 33, 255, 255, 255]]
    rows = hb.random_situations(1000, seed=20251019)
    assert rows.shape == (1000, 8) and rows.dtype == np.uint8
    assert (rows[:, 5:] == 255).all() and (rows[:, :5] < 52).all()
    assert all(len(set(r[:5])) == 5 for r in rows.tolist())
    assert np.array_equal(rows, hb.random_situations(1000, seed=20251019))
    with pytest.raises(ValueError):
        hb.pack_situations([[1, 2]], [[3, 4]])


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 9, 1000, 65536, 1000000):
        for w in (1, 2, 3, 4, 8):
            seen = []
            for r in range(w):
                lo, hi, per = hb.shard_bounds(n, w, r)
                assert 0 <= lo <= hi <= n and hi - lo <= per
                seen.extend(range(lo, hi)) if n <= 1000 else seen.append((lo, hi))
            if n <= 1000:
                assert seen == list(range(n))
            else:
                assert seen[0][0] == 0 and seen[-1][1] == n
                assert all(a[1] == b[0] for a, b in zip(seen, seen[1:]))


def test_float_formulas_match_oracle_formulas():
    from oracle import oracle as O
    rows = torch.tensor([[558, 397, 126, 334713, 148530, 172965, 70272, 265983, 130617, 27162, 16905, 104109,
                          558, 397, 126, 3],
                         [0, 0, 990, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 990, 5]], dtype=torch.int32)
    ppot, npot = hb.potentials_from_counts(rows)
    assert ppot.dtype == torch.float32
    assert float(ppot[0]) == 135.2781982421875 and float(npot[0]) == 327.26544189453125
    assert float(ppot[1]) == 0.0 and torch.isnan(npot[1])
    hs = hb.hs_from_counts(rows)
    assert float(hs[0]) == O.hs_float(558, 397, 126)
    pn, nn = hb.normalised_potentials(rows)
    assert abs(float(pn[0]) - 135.2781982421875 / 1176) < 1e-6
    ehs = hb.ehs_from_counts(rows)
    assert 0.0 <= float(ehs[0]) <= 1.0 and float(ehs[1]) == 0.0


def test_dropin_surface():
    import HandCalculations as HC
    import pokerlib_to_treys as P2T
    for name in ("HandStrength", "HandPotential", "CalculateHandRankValue", "CalculateHandRank", "evaluate_best_hand",
                 "remove_cards_from_deck", "generate_combinations", "generate_opponent_hands",
                 "generate_possible_turns_and_rivers", "is_straight", "determine_flush", "device", "evaluator",
                 "convert_to_treys", "CalculateHandValue", "Evaluator", "Card", "torch", "random", "Counter"):
        assert hasattr(HC, name), name
    assert HC.device in ("cuda", "cpu") and P2T.device in ("cuda:0", "cpu")
    pins = __import__("json").load(open(os.path.join(ROOT, "tests", "golden", "helper_pins.json")))
    assert len(HC.remove_cards_from_deck(torch.tensor([0, 1, 2]))) == pins["remove_3"]
    assert len(HC.remove_cards_from_deck(torch.tensor([52, 53]))) == pins["remove_unknown"]
    assert len(list(HC.generate_combinations(torch.tensor([0, 1, 2])))) == pins["comb_3"]
    assert len(list(HC.generate_opponent_hands(torch.tensor([0, 1]), torch.tensor([2, 3])))) == pins["opp_4known"]
    assert len(list(HC.generate_possible_turns_and_rivers(torch.tensor([0, 1, 2])))) == pins["runouts_flop"]
    first = [t.tolist() for t in list(HC.generate_opponent_hands(torch.tensor([14, 1]), torch.tensor([31, 32, 33])))[:5]]
    assert first == pins["first_opp_hands"]
    for values, expect in pins["is_straight"]:
        assert HC.is_straight(values) == expect
    assert HC.determine_flush([0] * 5, {0: 5}) == (True, 0) and HC.determine_flush([], {1: 4}) == (False, None)
    assert HC.CalculateHandRank([0, 1], [2, 3]) == (0,)
    with pytest.raises(TypeError):
        HC.CalculateHandRank([0, 1], [2, 3, 4])

    class _R:  # duck-typed poker.Card
        def __init__(self, s):
            self.rank = type("R", (), {"val": s[0]})()
            self.suit = type("S", (), {"value": ("x", s[1])})()

    assert P2T.convert_to_treys([_R("Kd"), _R("5s")]) == [0x08004B25, 0x00081307]
    assert P2T.Card.new("Jc") == 0x0200891D and P2T.Card.int_to_str(0x0200891D) == "Jc"
    ev = P2T.Evaluator()
    assert ev.get_rank_class(1) == 1 and ev.get_rank_class(1600) == 5 and ev.get_rank_class(7462) == 9
    assert ev.class_to_string(5) == "Straight"
    assert HC.get_mode() == os.environ.get("HOLDEM_MODE", "reference")    # drop-in default: the reference's own semantics
    HC.set_mode("treys")
    assert HC.get_mode() == "treys"
    HC.set_mode("reference")
    with pytest.raises(ValueError):
        HC.set_mode("bogus")


def test_policy_matches_reference_decision_golden():
    """Batched Decision mirror vs golden vectors from the unmodified reference Decision.py (tests/golden/)."""
    import json
    from holdem_b200 import policy
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "decision_golden.json")))["cases"]
    for pos in ("SB", "BB", "NONE"):
        sub = [c for c in g if c[3] == pos]
        hs = torch.tensor([c[0] for c in sub], dtype=torch.float64)
        pp = torch.tensor([c[1] for c in sub], dtype=torch.float64)
        npot = torch.tensor([c[2] for c in sub], dtype=torch.float64)
        codes = policy.decide_batch(hs, pp, npot, pos).tolist()
        want = [policy.NONE_TAKEN if c[4] is None else policy.ACTIONS.index(c[4]) for c in sub]
        assert codes == want
    # mixed positions in one call + the scalar form with the reference's signature
    hs = torch.tensor([c[0] for c in g], dtype=torch.float64)
    pp = torch.tensor([c[1] for c in g], dtype=torch.float64)
    npot = torch.tensor([c[2] for c in g], dtype=torch.float64)
    pos = torch.tensor([policy.POSITIONS[c[3]] for c in g])
    assert policy.decide_batch(hs, pp, npot, pos).tolist() == \
        [policy.NONE_TAKEN if c[4] is None else policy.ACTIONS.index(c[4]) for c in g]
    for c in g[::97]:
        assert policy.decision_based_on_blind_position(c[0], c[1], c[2], c[3]) == c[4]
    with pytest.raises(ValueError):
        policy.decision_based_on_blind_position(0.5, 0.5, 0.5, "UTG")
    rows = torch.tensor([[558, 397, 126, 334713, 148530, 172965, 70272, 265983, 130617, 27162, 16905, 104109,
                          558, 397, 126, 3]], dtype=torch.int32)
    assert policy.decide_from_counts(rows, "BB").tolist() == [policy.CHECK]


def test_table_helpers_cpu():
    from holdem_b200 import table
    hole, board = table.deal_tables(50, seed=20251020, seats=22)
    assert hole.shape == (50, 22, 2) and board.shape == (50, 5)
    for h in range(50):
        cards = hole[h].reshape(-1).tolist() + board[h].tolist()
        assert len(set(cards)) == 49 and max(cards) < 52
    sit = table.situations_for_street(hole, board, 4)
    assert sit.shape == (1100, 8) and (sit[:, 6:] == 255).all() and (sit[:, 2:6] == np.repeat(board[:, :4], 22, 0)).all()
    ranks = torch.tensor([[10, 10, 7000, 3], [5, 9, 9, 2]], dtype=torch.int16)
    assert table.winners(ranks, "best").tolist() == [[False, False, False, True], [False, False, False, True]]
    assert table.winners(ranks, "reference").tolist() == [[False, False, True, False], [False, True, True, False]]
    active = torch.tensor([[True, True, True, False], [True, True, True, True]])
    assert table.winners(ranks, "best", active).tolist() == [[True, True, False, False], [False, False, False, True]]


def test_torch_ops_library_loads_and_has_no_cpu_kernel():
    """torch.ops.holdem.* (csrc/torch_ops.cpp) registers without a GPU; only CUDA implementations exist."""
    from holdem_b200 import _ops
    ops = _ops.load()
    for name in ("eval", "category", "hs", "hp", "hp_out", "hp_peers", "hp_sharded", "table_hp", "decide",
                 "decide_from_counts", "pack_rank_major", "rows_hash"):
        assert hasattr(ops, name), name
    with pytest.raises(Exception) as ei:
        ops.hp(torch.zeros((1, 8), dtype=torch.uint8), 0)
    assert "CPU" in str(ei.value)


def test_rows_hash_host_properties():
    rng = np.random.default_rng(0)
    rows = rng.integers(0, 2 ** 31, size=(1000, 16), dtype=np.int64).astype(np.int32)
    h = hb.rows_hash_host(rows)
    assert h == (hb.rows_hash_host(rows[:400]) + hb.rows_hash_host(rows[400:], first_row=400)) % 2 ** 64
    swapped = rows.copy(); swapped[[1, 2]] = swapped[[2, 1]]
    assert hb.rows_hash_host(swapped) != h and hb.rows_hash_host(rows, first_row=1) != h
    assert hb.rows_hash_host(rows[:0]) == 0


def test_positions_and_tensor_converters_cpu():
    from holdem_b200 import policy
    assert [policy.position_from_player_position(p) for p in
            ("NONE", "SMALL_BLIND", "BIG_BLIND", "DEALER", "DEALER_SMALL_BLIND")] == ["NONE", "SB", "BB", "NONE", "SB"]
    assert [policy.position_from_player_position(v) for v in (1, 2, 3, 4, 5)] == ["NONE", "SB", "BB", "NONE", "SB"]

    class P:            # duck-typed Enum member (utils.PlayerPosition needs the third-party `poker` import chain)
        name = "BIG_BLIND"
    assert policy.position_from_player_position(P()) == "BB"
    with pytest.raises(ValueError):
        policy.position_from_player_position("UTG")
    assert policy.position_codes(["SB", "BIG_BLIND", "DEALER"]).tolist() == [0, 1, 2]
    c = torch.arange(52)
    assert hb.rank_major_from_path_t(c).tolist() == [hb.rank_major_from_path(int(x)) for x in c]
    assert torch.equal(hb.path_from_rank_major_t(hb.rank_major_from_path_t(c)), c)
    assert hb.treys_from_path_t(c).tolist() == [hb.treys_from_path(int(x)) for x in c]
    assert torch.equal(hb.path_from_treys_t(hb.treys_from_path_t(c)), c)
    with pytest.raises(ValueError):
        hb.path_from_treys_t(torch.tensor([12345]))
    rows = hb.situations_from_rank_major(torch.tensor([[0, 5]]), torch.tensor([[51, 50, 4]]))
    assert rows.tolist() == [[0, 14, 51, 38, 1, 255, 255, 255]]
    with pytest.raises(ValueError):
        hb.pack_situations([[0, 300]], [[1, 2, 3]])
    with pytest.raises(ValueError):
        hb.pack_situations([[0, -1]], [[1, 2, 3]])


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the reference's algorithm on the host cores) needs no GPU: one JSON line with the contract's keys
    and the same `config` as our arm."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "flop_hs_hp_situations_per_sec" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert d["config"] == mod.CONFIG


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_bench_our_arm_fails_loudly_without_a_gpu():
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
