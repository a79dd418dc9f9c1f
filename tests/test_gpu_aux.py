"""GPU tests of the pieces either side of the HS/HP kernels: the torch.ops.holdem layer against the
ctypes binding, the content hash, the fused decision kernel against the reference's Decision.py golden cases, the
wire-format packing kernel, whole-table evaluation, the explicit *_variant entry points and the per-call flag slots."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _mixed(hb, n=900):
    parts = [hb.random_situations(n // 3, seed=5), hb.random_situations(n // 3, seed=6, n_board=4),
             hb.random_situations(n // 3, seed=7, n_board=5)]
    sit = np.concatenate(parts)
    return sit[np.random.default_rng(3).permutation(len(sit))]


def test_torch_ops_equal_ctypes_binding(hb):
    """Both bindings of the C ABI return the same tensors: torch.ops.holdem.* (csrc/torch_ops.cpp) and ctypes (_lib.py)."""
    for name in ("eval", "hs", "hp", "hp_out", "hp_peers", "hp_sharded", "table_hp", "decide_from_counts"):
        assert hasattr(torch.ops.holdem, name), name
    sit = torch.from_numpy(_mixed(hb)).cuda()
    for mode in ("treys", "reference"):
        m = hb.mode_id(mode)
        sub = sit if mode == "treys" else sit[:96].contiguous()
        assert torch.equal(torch.ops.holdem.hp(sub, m), hb.hand_potential_counts(sub, mode, binding="ctypes"))
        assert torch.equal(torch.ops.holdem.hs(sub, m), hb.hand_strength_counts(sub, mode, binding="ctypes"))
        out = torch.zeros((sub.shape[0], 16), dtype=torch.int32, device="cuda")
        torch.ops.holdem.hp_out(sub, m, out)
        assert torch.equal(out, hb.hand_potential_counts(sub, mode))
    hands = torch.from_numpy(hb.random_situations(5000, seed=8, n_board=5)).cuda()
    for nc in (5, 6, 7):
        assert torch.equal(torch.ops.holdem.eval(hands, nc, False), hb.evaluate(hands, nc, binding="ctypes"))
        assert torch.equal(torch.ops.holdem.eval(hands, nc, True), hb.evaluate(hands, nc, strict=True, binding="ctypes"))
    assert torch.equal(torch.ops.holdem.category(hands, 7), hb.category(hands, 7, binding="ctypes"))
    # no CPU implementation is registered, and shapes / dtypes are checked
    with pytest.raises(Exception):
        torch.ops.holdem.hp(sit.cpu(), 0)
    with pytest.raises(Exception):
        torch.ops.holdem.hp(sit[:, :7].contiguous(), 0)
    with pytest.raises(Exception):
        torch.ops.holdem.hp(sit, 7)
    # the ops run on torch's CURRENT stream
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        a = torch.ops.holdem.hp(sit, 0)
    s.synchronize()
    assert torch.equal(a, hb.hand_potential_counts(sit, "treys"))


def test_current_device_is_restored(hb):
    cur = torch.cuda.current_device()
    sit = torch.from_numpy(hb.random_situations(8, seed=1)).cuda()
    hb.hand_potential_counts(sit, "treys", binding="ctypes")
    hb.hand_potential_counts_host(hb.random_situations(8, seed=1), "treys", device=0)
    assert torch.cuda.current_device() == cur


def test_rows_hash(hb):
    sit = torch.from_numpy(_mixed(hb, 3000)).cuda()
    rows = hb.hand_potential_counts(sit, "treys")
    h = hb.rows_hash(rows)
    assert h == hb.rows_hash_host(rows.cpu().numpy())                       # device kernel == numpy twin
    assert h == int(torch.ops.holdem.rows_hash(rows, 0).item()) & (2 ** 64 - 1)
    assert h == (hb.rows_hash(rows[:1000].contiguous()) + hb.rows_hash(rows[1000:].contiguous(), first_row=1000)) % 2 ** 64
    # sensitive to a single changed count, to a swap of two rows, and to a shift of the block
    r2 = rows.clone(); r2[1234, 7] += 1
    assert hb.rows_hash(r2) != h
    r3 = rows.clone(); r3[[10, 11]] = r3[[11, 10]]
    assert (not torch.equal(r3[10], r3[11])) and hb.rows_hash(r3) != h
    assert hb.rows_hash(rows, first_row=1) != h
    assert hb.rows_hash(rows[:0].contiguous()) == 0
    # the sum of all counts, by contrast, cannot tell flop rows apart ("checksum_of_counts is a structural constant")
    flop = hb.hand_potential_counts(torch.from_numpy(hb.random_situations(64, seed=2)).cuda(), "treys")
    assert len(set(flop[:, :15].sum(dim=1).tolist())) == 1


def test_decide_kernels_match_reference_golden(hb):
    """The 2,940 cases produced by the unmodified Decision.py through the fused kernel, per-row positions and whole-batch
    positions; and decide_from_counts against the torch expressions on real count rows of every street."""
    from holdem_b200 import policy
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "decision_golden.json")))["cases"]
    hs = torch.tensor([c[0] for c in g], dtype=torch.float64, device="cuda")
    pp = torch.tensor([c[1] for c in g], dtype=torch.float64, device="cuda")
    npot = torch.tensor([c[2] for c in g], dtype=torch.float64, device="cuda")
    pos = policy.position_codes([c[3] for c in g], device="cuda")
    want = [policy.NONE_TAKEN if c[4] is None else policy.ACTIONS.index(c[4]) for c in g]
    got = policy.decide_batch(hs, pp, npot, pos)
    assert got.dtype == torch.int8 and got.is_cuda and got.tolist() == want
    for name, code in policy.POSITIONS.items():
        sel = pos == code
        assert policy.decide_batch(hs[sel], pp[sel], npot[sel], name).tolist() == [w for w, s in zip(want, sel.tolist()) if s]
    assert torch.equal(policy.decide_batch(hs.cpu(), pp.cpu(), npot.cpu(), pos.cpu()), got.cpu())     # CPU expressions agree
    with pytest.raises(ValueError):
        policy.decide_batch(hs, pp, npot, "UTG")
    # from count rows: one fused pass == hs_from_counts + normalised_potentials + the tree
    sit = torch.from_numpy(_mixed(hb, 3000)).cuda()
    rows = hb.hand_potential_counts(sit, "treys")
    posr = torch.randint(0, 3, (rows.shape[0],), dtype=torch.uint8, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
    fused = policy.decide_from_counts(rows, posr)
    hs2 = hb.hs_from_counts(rows)
    pp2, np2 = hb.normalised_potentials(rows)
    ref = policy.decide_batch(hs2.cpu(), torch.nan_to_num(pp2, nan=0.0).cpu(), torch.nan_to_num(np2, nan=0.0).cpu(), posr.cpu())
    assert torch.equal(fused.cpu(), ref)
    assert len(set(fused.tolist())) >= 3
    # the three doubles the kernel derives are bit-identical to the torch float64 expressions
    trip = torch.empty((rows.shape[0], 3), dtype=torch.float64, device="cuda")
    act = torch.empty(rows.shape[0], dtype=torch.int8, device="cuda")
    from holdem_b200 import _lib
    from holdem_b200.api import _enter
    with _enter(rows) as (lib, st):
        _lib.check(lib.holdem_decide_from_counts(rows.data_ptr(), posr.data_ptr(), 2, rows.shape[0], act.data_ptr(),
                                                 trip.data_ptr(), st))
    assert torch.equal(act, fused)
    assert torch.equal(trip[:, 0], hs2)
    assert torch.equal(trip[:, 1], torch.nan_to_num(pp2, nan=0.0)) and torch.equal(trip[:, 2], torch.nan_to_num(np2, nan=0.0))
    # a malformed row gives -2
    bad = rows.clone(); bad[3] = -1
    assert int(policy.decide_from_counts(bad, "BB")[3]) == -2
    assert policy.position_from_player_position("DEALER_SMALL_BLIND") == "SB" and policy.position_from_player_position(3) == "BB"


def test_pack_rank_major_kernel(hb, oracle):
    """Wire ints (rank*4+suit, utils.py:82-85) -> situation rows on the device; same rows as the host packer, and the HS/HP
    kernels give the oracle's counts on them."""
    sit = hb.random_situations(500, seed=17, n_board=4)
    hole_rm = torch.tensor([[hb.rank_major_from_path(c) for c in r[:2]] for r in sit], dtype=torch.int64)
    board_rm = torch.tensor([[hb.rank_major_from_path(c) for c in r[2:6]] for r in sit], dtype=torch.int64)
    rows = hb.situations_from_rank_major(hole_rm.cuda(), board_rm.cuda())
    assert rows.dtype == torch.uint8 and np.array_equal(rows.cpu().numpy(), sit)
    assert torch.equal(hb.situations_from_rank_major(hole_rm, board_rm), rows.cpu())          # CPU packer agrees
    assert np.array_equal(hb.hand_potential_counts(rows, "treys").cpu().numpy().astype(np.uint32),
                          oracle.hp_batch(sit, "treys"))
    bad_hole = hole_rm.clone(); bad_hole[4, 0] = 52; bad_hole[9, 1] = -1
    rows2 = hb.situations_from_rank_major(bad_hole.cuda(), board_rm.cuda())
    out = hb.hand_potential_counts(rows2, "treys")
    assert bool((out[4] == -1).all()) and bool((out[9] == -1).all()) and bool((out[5] != -1).all())
    with pytest.raises(ValueError):
        hb.situations_from_rank_major(bad_hole, board_rm)
    with pytest.raises(ValueError):
        hb.pack_situations([[0, 256]], [[1, 2, 3]])                                       # no silent uint8 wrap
    # vectorised converters round-trip
    c = torch.arange(52, device="cuda")
    assert torch.equal(hb.path_from_rank_major_t(hb.rank_major_from_path_t(c)), c)
    assert torch.equal(hb.path_from_treys_t(hb.treys_from_path_t(c)), c)
    assert hb.treys_from_path_t(torch.tensor([hb.path_from_str("Kd")])).tolist() == [0x08004B25]


def test_table_hp_matches_row_expansion(hb, oracle):
    """holdem_table_hp_batch (22 seats share a board, GameBoard.py:12) == holdem_hp_batch on the expanded rows, every street,
    both modes; a sample of the rows against the oracle."""
    from holdem_b200 import table
    hole, board = table.deal_tables(40, seed=20251020, seats=22)
    d_hole, d_board = torch.from_numpy(hole).cuda(), torch.from_numpy(board).cuda()
    for nb in (3, 4, 5):
        rows = table.situations_for_street(hole, board, nb)
        want = hb.hand_potential_counts(torch.from_numpy(rows).cuda(), "treys")
        got = hb.table_hand_potential_counts(d_hole, d_board, nb, "treys")
        assert got.shape == (40, 22, 16) and torch.equal(got.reshape(-1, 16), want)
        assert torch.equal(hb.table_hand_potential_counts(d_hole, d_board, nb, "treys", binding="ctypes"), got)
        pick = np.random.default_rng(nb).choice(len(rows), 24 if nb == 3 else 64, replace=False)
        assert np.array_equal(got.reshape(-1, 16)[torch.from_numpy(pick).cuda()].cpu().numpy().astype(np.uint32),
                              oracle.hp_batch(rows[pick], "treys", cache4=True))
    small_h, small_b = d_hole[:2, :3].contiguous(), d_board[:2]
    rows = table.situations_for_street(hole[:2, :3], board[:2], 3)
    assert torch.equal(hb.table_hand_potential_counts(small_h, small_b, 3, "reference").reshape(-1, 16),
                       hb.hand_potential_counts(torch.from_numpy(rows).cuda(), "reference"))
    with pytest.raises(Exception):
        hb.table_hand_potential_counts(d_hole, d_board, 2, "treys")


def test_variant_entry_points(hb):
    """The A/B kernels are reached through explicit entry points (no environment switches in the production dispatch)."""
    sit = torch.from_numpy(_mixed(hb, 9000)).cuda()       # >= 4096 rows: the flop kernel works them in texture order
    want = hb.hand_potential_counts(sit, "treys")
    flop = torch.from_numpy(hb.random_situations(8192, seed=77)).cuda()
    assert torch.equal(hb.hand_potential_counts(flop, "treys", flop_variant=5), hb.hand_potential_counts(flop, "treys", flop_variant=4))
    assert torch.equal(hb.hand_potential_counts(flop, "treys", flop_variant=6), hb.hand_potential_counts(flop, "treys", flop_variant=4))
    assert torch.equal(hb.hand_potential_counts(flop, "treys", flop_variant=7), hb.hand_potential_counts(flop, "treys", flop_variant=4))
    for variants in ((4, 4, 0), (3, 3, 1), (2, 4, 1), (4, 3, 0), (5, 4, 0), (6, 4, 0), (7, 4, 0)):
        assert torch.equal(hb.hand_potential_counts(sit, "treys", variants=variants), want), variants
    for mode in ("treys", "reference"):
        assert torch.equal(hb.hand_strength_counts(sit, mode, variant=1), hb.hand_strength_counts(sit, mode))
    h0, s0 = hb.exhaustive7(0)
    h1, s1 = hb.exhaustive7(0, variant=1)
    assert torch.equal(h0, h1) and torch.equal(s0, s1)
    pin_in = torch.from_numpy(_mixed(hb, 300)).pin_memory()
    a = hb.hand_potential_counts_host(pin_in.numpy(), "treys")
    b = hb.hand_potential_counts_host(pin_in.numpy(), "treys", staged=True)
    assert np.array_equal(a, b)
    src = open(os.path.join(ROOT, "dask-holdem-simulator_b200", "csrc", "capi.cu")).read()
    assert "getenv" not in src


def test_flag_slots_many_calls_in_flight(hb):
    """More HP calls than the 1,024-slot ring holds, queued on four streams without synchronising: every call must still see
    its own street flags (a reused slot waits for the call that used it last)."""
    flop = torch.from_numpy(hb.random_situations(6, seed=1)).cuda()
    mixed = torch.from_numpy(_mixed(hb, 30)).cuda()
    want_flop, want_mixed = hb.hand_potential_counts(flop, "treys"), hb.hand_potential_counts(mixed, "treys")
    streams = [torch.cuda.Stream() for _ in range(4)]
    outs = []
    torch.cuda.synchronize()
    for i in range(2600):
        with torch.cuda.stream(streams[i % 4]):
            src = mixed if i % 3 == 0 else flop
            o = torch.full((src.shape[0], 16), -7, dtype=torch.int32, device="cuda")
            hb.hand_potential_counts(src, "treys", out=o)
            if i % 50 == 0 or i > 2500:
                outs.append((i, o))
    torch.cuda.synchronize()
    for i, o in outs:
        assert torch.equal(o, want_mixed if i % 3 == 0 else want_flop), i
    assert hb.kernel_launches() > 2600 * 3
