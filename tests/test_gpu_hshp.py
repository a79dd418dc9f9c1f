"""GPU parity: hand strength and hand potential counts vs the CPU oracle and vs the golden vectors produced by the
unmodified reference, through the C ABI; plus size-independent properties at BASELINE.json's full sizes."""
import struct

import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def f32hex(x):
    return struct.pack(">f", float(x)).hex()


def _u32(t):
    return t.cpu().numpy().astype(np.uint32)


@pytest.mark.parametrize("mode", ["treys", "reference"])
@pytest.mark.parametrize("n_board", [3, 4, 5])
def test_hs_matches_oracle(hb, oracle, mode, n_board):
    sit = hb.random_situations(600, seed=40 + n_board, n_board=n_board)
    got = _u32(hb.hand_strength_counts(torch.from_numpy(sit).cuda(), mode))
    assert np.array_equal(got, oracle.hs_batch(sit, mode))
    assert np.array_equal(hb.hand_strength_counts_host(sit, mode), got)


def test_hs_reference_golden_floats(hb):
    cases = load_golden("literal_hs.json")["cases"]
    for nb in (3, 4, 5):
        sub = [c for c in cases if len(c["board"]) == nb]
        sit = hb.pack_situations([c["our"] for c in sub], [c["board"] for c in sub])
        counts = hb.hand_strength_counts(torch.from_numpy(sit).cuda(), "reference")
        hs = hb.hs_from_counts(counts).cpu().tolist()
        for c, v, cnt in zip(sub, hs, counts.cpu().tolist()):
            assert float(v).hex() == c["hs_hex"], c
            a, t, b = cnt
            assert ((a + t / 2) / (a + t + b)).hex() == c["hs_hex"]


def test_hp_reference_golden(hb):
    """HP[3][3], HPTotal[3] and the float32 PPot/NPot bit-identical to the unmodified reference (flop/turn/river)."""
    cases = load_golden("literal_hp.json")["cases"]
    for nb in (3, 4, 5):
        sub = [c for c in cases if len(c["board"]) == nb]
        sit = hb.pack_situations([c["our"] for c in sub], [c["board"] for c in sub])
        rows = hb.hand_potential_counts(torch.from_numpy(sit).cuda(), "reference")
        ppot, npot = hb.potentials_from_counts(rows)
        assert ppot.dtype == torch.float32 and ppot.is_cuda
        for c, row, p, q in zip(sub, rows.cpu().tolist(), ppot.cpu().tolist(), npot.cpu().tolist()):
            assert [row[3:6], row[6:9], row[9:12]] == c["HP"], c
            assert row[12:15] == c["HPTotal"] and row[0:3] == c["HPTotal"] and row[15] == nb
            assert f32hex(p) == c["ppot_f32_hex"] and f32hex(q) == c["npot_f32_hex"]


@pytest.mark.parametrize("mode,n_board,n", [("treys", 3, 48), ("treys", 4, 200), ("treys", 5, 200),
                                            ("reference", 3, 48), ("reference", 4, 24), ("reference", 5, 24)])
def test_hp_matches_oracle(hb, oracle, mode, n_board, n):
    sit = hb.random_situations(n, seed=60 + n_board, n_board=n_board)
    want = oracle.hp_batch(sit, mode)
    d_sit = torch.from_numpy(sit).cuda()
    assert np.array_equal(_u32(hb.hand_potential_counts(d_sit, mode)), want)
    assert np.array_equal(_u32(hb.hand_potential_counts(d_sit, mode, direct=True)), want)


def test_hp_treys_regression_vectors(hb):
    sit = hb.pack_situations([[14, 1]], [[31, 32, 33]])
    row = hb.hand_potential_counts(torch.from_numpy(sit).cuda(), "treys")[0].cpu().tolist()
    assert row[0:3] == [564, 1, 516]
    assert row[3:12] == [276753, 35790, 245817, 0, 658, 332, 28713, 33025, 449102]
    ppot, npot = hb.potentials_from_counts(torch.tensor([row], dtype=torch.int32))
    assert float(ppot[0]) == 87.47679138183594 and float(npot[0]) == 467.04071044921875


def test_hp_host_entry_and_bad_rows(hb, oracle):
    sit = hb.random_situations(40, seed=77)
    assert np.array_equal(hb.hand_potential_counts_host(sit, "treys"), oracle.hp_batch(sit, "treys"))
    bad = sit.copy()
    bad[3, 1] = bad[3, 0]            # duplicate hole card
    bad[5, 2] = 99                   # unknown card
    bad[7, 2:7] = [1, 255, 2, 255, 255]  # hole in the board
    bad[9, 2:7] = [1, 2, 255, 255, 255]  # two-card board
    rows = hb.hand_potential_counts(torch.from_numpy(bad).cuda(), "treys").cpu().numpy()
    hs = hb.hand_strength_counts(torch.from_numpy(bad).cuda(), "treys").cpu().numpy()
    for i in (3, 5, 7, 9):
        assert (rows[i] == -1).all() and (hs[i] == -1).all(), i
    good = [i for i in range(40) if i not in (3, 5, 7, 9)]
    assert np.array_equal(rows[good].astype(np.uint32), oracle.hp_batch(sit[good], "treys"))
    with pytest.raises(hb.HoldemError):
        hb.hand_potential_counts_host(bad, "treys")
    assert hb.hand_potential_counts(torch.empty((0, 8), dtype=torch.uint8, device="cuda")).shape == (0, 16)
    with pytest.raises(ValueError):
        hb.hand_potential_counts(torch.from_numpy(sit).cuda(), "bogus")


def test_hp_full_size_properties(hb):
    """BASELINE.json configs[2]: 65,536 random flop situations.  Size-independent checks on every row plus equality of
    the two independent GPU implementations (deduplicating vs plain enumeration) on a 4,096-row slice."""
    n = 65536
    sit = hb.random_situations(n, seed=20251019)
    d_sit = torch.from_numpy(sit).cuda()
    rows = hb.hand_potential_counts(d_sit, "treys")
    r = rows.cpu().numpy().astype(np.int64)
    hp = r[:, 3:12].reshape(n, 3, 3)
    assert (r[:, 15] == 3).all()
    assert (r[:, 12:15].sum(axis=1) == 1081).all()
    assert (r[:, 0:3] == r[:, 12:15]).all()
    assert (hp.sum(axis=2) == 990 * r[:, 12:15]).all()        # every opponent holding sees exactly C(45,2) runouts
    hs = hb.hand_strength_counts(d_sit, "treys").cpu().numpy()
    assert np.array_equal(hs, r[:, 0:3])                         # HS kernel agrees with the HP kernel's by-product
    again = hb.hand_potential_counts(d_sit, "treys")
    assert torch.equal(rows, again)                              # deterministic
    sl = slice(0, 4096)
    direct = hb.hand_potential_counts(d_sit[sl].contiguous(), "treys", direct=True)
    assert torch.equal(rows[sl], direct)


def test_hp_reference_v5_kernel_full_size(hb, oracle):
    """mode=reference (HandCalculations.py:178-226 as written): the rank-multiset kernel (kernels_hp_ref5.cu) against the plain
    enumeration kernel -- one literal category evaluation per (holding, runout) pair -- on 65,536 flop + 65,536 turn + 16,384
    river situations, against the earlier 4-set kernel, and against the C oracle on board textures that stress the flush part
    (monotone / four- and five-suited boards, paired boards, our own flush draws)."""
    textures = [([14, 1], [31, 32, 33]), ([0, 1], [2, 3, 4]), ([13, 26], [2, 3, 4]), ([5, 7], [1, 3, 9]), ([5, 20], [1, 3, 30]),
                ([12, 25], [0, 14, 28]), ([12, 25], [11, 24, 37]), ([0, 13], [26, 39, 1]), ([24, 48], [26, 2, 16]),
                ([0, 1], [2, 3, 4, 5]), ([13, 26], [2, 3, 4, 5]), ([12, 11], [2, 3, 4, 18]), ([39, 1], [41, 42, 43, 44]),
                ([0, 12], [13, 26, 39, 1]), ([21, 47], [8, 34, 9, 35]),
                ([0, 1], [2, 3, 4, 5, 6]), ([13, 26], [2, 3, 4, 5, 7]), ([12, 25], [0, 2, 4, 6, 8]), ([11, 13], [0, 2, 4, 6, 21]),
                ([0, 13], [26, 27, 28, 29, 30]), ([5, 18], [31, 44, 6, 19, 32]), ([50, 51], [49, 48, 47, 46, 45])]
    for nb in (3, 4, 5):
        sub = [c for c in textures if len(c[1]) == nb]
        sit = hb.pack_situations([c[0] for c in sub], [c[1] for c in sub])
        want = oracle.hp_batch(sit, "reference")
        d = torch.from_numpy(sit).cuda()
        for variant in (5, 4):
            got = _u32(hb.hand_potential_counts(d, "reference", ref_variant=variant))
            bad = np.nonzero((got != want).any(axis=1))[0]
            assert bad.size == 0, (nb, variant, bad, sit[bad[:3]], got[bad[:3]], want[bad[:3]])
    for nb, n in ((3, 65536), (4, 65536), (5, 16384)):
        d = torch.from_numpy(hb.random_situations(n, seed=700 + nb, n_board=nb)).cuda()
        v5 = hb.hand_potential_counts(d, "reference")
        direct = hb.hand_potential_counts(d, "reference", direct=True)
        bad = torch.nonzero((v5 != direct).any(dim=1)).flatten()
        assert bad.numel() == 0, (nb, bad[:5].tolist(), d[bad[:2]].tolist(), v5[bad[:2]].tolist(), direct[bad[:2]].tolist())
        sl = slice(0, 4096)
        assert torch.equal(hb.hand_potential_counts(d[sl].contiguous(), "reference", ref_variant=4), v5[sl])
        hs = hb.hand_strength_counts(d, "reference")
        assert torch.equal(hs, v5[:, 0:3])


def test_hp_treys_all_65536_vs_oracle(hb, oracle):
    """EVERY row of the headline workload (BASELINE.json configs[2]: 65,536 random flop situations, seed 20251019)
    against the C oracle -- made affordable by the oracle's 4-set cache (oracle/ehs_oracle.c:
    ehs_hp_counts_treys_4set, itself checked against the plain loop in tests/test_oracle_ehs.py), all host cores."""
    n = 65536
    sit = hb.random_situations(n, seed=20251019)
    got = _u32(hb.hand_potential_counts(torch.from_numpy(sit).cuda(), "treys"))
    want = oracle.hp_batch_parallel(sit, "treys", cache4=True)
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert bad.size == 0, (bad[:8], got[bad[:2]], want[bad[:2]])


def test_dropin_handcalculations(hb):
    import HandCalculations as HC
    our_hs, board_hs = torch.tensor([0, 13]), torch.tensor([26, 27, 28, 29, 30])     # testing/ExampleHS.py:4-5
    our_hp, board_hp = torch.tensor([14, 1]), torch.tensor([31, 32, 33])            # testing/ExampleHS.py:11-12
    assert HC.get_mode() == "reference"            # the drop-in's default: what the original module returns
    try:
        hs = HC.HandStrength(our_hs, board_hs)
        assert isinstance(hs, float) and hs == 0.5
        hp = HC.HandPotential(our_hp, board_hp)
        assert isinstance(hp, list) and len(hp) == 2
        assert all(isinstance(x, torch.Tensor) and x.dim() == 0 and x.dtype == torch.float32 and x.is_cuda for x in hp)
        assert [float(x) for x in hp] == [135.2781982421875, 327.26544189453125]
        s = HC.HandStrength(torch.tensor([0, 1]), torch.tensor([2, 3, 4]))             # test_HandCalculations.py:49-52
        assert s == 0.9995374653098983 and 0 <= s <= 1
        assert our_hs.tolist() == [0, 13]                                              # inputs are not mutated
        with pytest.raises(ValueError):
            HC.HandPotentialCounts(torch.tensor([0, 0]), board_hp)                     # malformed rows raise everywhere
        with pytest.raises(ValueError):
            HC.EffectiveHandStrength(torch.tensor([0, 31]), board_hp)
        HC.set_mode("treys")
        assert HC.HandStrength(our_hs, board_hs) == (0 + 946 / 2) / 990
        hp = HC.HandPotential(our_hp, board_hp)
        assert [float(x) for x in hp] == [87.47679138183594, 467.04071044921875]
        assert HC.HandPotentialCounts(our_hp, board_hp)[1] == [564, 1, 516]
        assert 0.0 <= HC.EffectiveHandStrength(our_hp, board_hp) <= 1.0
    finally:
        HC.set_mode("reference")
    assert HC.CalculateHandRankValue(torch.tensor([0, 1]), torch.tensor([10, 11, 12, 13, 14])) == 8
    assert HC.CalculateHandRankValue([0, 1], [10, 11, 12, 13, 14]) == 8                # test_HandCalculations.py:17-18

    class _C:  # duck-typed poker.Card
        def __init__(self, s):
            self.rank = type("R", (), {"val": s[0]})()
            self.suit = type("S", (), {"value": ("x", s[1])})()

    assert HC.evaluate_best_hand([_C("Qs"), _C("Th")], [_C("Ah"), _C("Kd"), _C("Jc")]) == 1600
    assert HC.evaluate_best_hand([_C("As"), _C("Ks")], [_C("Qs"), _C("Js"), _C("Ts"), _C("2c"), _C("2d")]) == 1
    import pokerlib_to_treys as P2T
    assert P2T.CalculateHandValue([_C("Ah"), _C("Kd"), _C("Jc")], [_C("Qs"), _C("Th")]) == 1 - 1600 / 7462
    with pytest.raises(KeyError):
        HC.evaluate_best_hand([_C("Qs"), _C("Th")], [_C("Ah"), _C("Kd")])
    with pytest.raises(ValueError):
        HC.HandStrength(torch.tensor([0, 0]), torch.tensor([2, 3, 4]))


def test_hp_mixed_streets_in_one_batch(hb, oracle):
    """Flop, turn and river rows interleaved in one call: the three kernels each take their rows."""
    parts = [hb.random_situations(40, seed=90 + nb, n_board=nb) for nb in (3, 4, 5)]
    sit = np.concatenate(parts)
    rng = np.random.default_rng(3)
    sit = sit[rng.permutation(sit.shape[0])]
    for mode in ("treys", "reference"):
        got = _u32(hb.hand_potential_counts(torch.from_numpy(sit).cuda(), mode))
        assert np.array_equal(got, oracle.hp_batch(sit, mode)), mode
    only_turn = parts[1]
    assert np.array_equal(_u32(hb.hand_potential_counts(torch.from_numpy(only_turn).cuda(), "treys")),
                          oracle.hp_batch(only_turn, "treys"))
    only_river = parts[2]
    assert np.array_equal(_u32(hb.hand_potential_counts(torch.from_numpy(only_river).cuda(), "treys")),
                          oracle.hp_batch(only_river, "treys"))


def test_hp_turn_full_size_properties(hb):
    """65,536 random turn situations: size-independent checks + equality with the plain-enumeration kernel."""
    n = 65536
    sit = hb.random_situations(n, seed=20251022, n_board=4)
    d_sit = torch.from_numpy(sit).cuda()
    rows = hb.hand_potential_counts(d_sit, "treys")
    r = rows.cpu().numpy().astype(np.int64)
    assert (r[:, 15] == 4).all() and (r[:, 12:15].sum(axis=1) == 1035).all()
    assert (r[:, 3:12].reshape(n, 3, 3).sum(axis=2) == 44 * r[:, 12:15]).all()
    assert np.array_equal(hb.hand_strength_counts(d_sit, "treys").cpu().numpy(), r[:, 0:3])
    direct = hb.hand_potential_counts(d_sit[:8192].contiguous(), "treys", direct=True)
    assert torch.equal(rows[:8192], direct)


def _oracle_chunk(args):
    from oracle import oracle as O
    sit_bytes, n, mode = args
    sit = np.frombuffer(sit_bytes, dtype=np.uint8).reshape(n, 8)
    return O.hp_batch(sit, mode).tobytes()


def test_hp_treys_1024_flops_vs_oracle(hb):
    """SURVEY 8d config 3 parity sample: the first 1,024 situations of the 65,536-situation bench workload, every count
    of every row against the C oracle (run on all host cores)."""
    import multiprocessing as mp
    import os
    sit = hb.random_situations(65536, seed=20251019)[:1024]
    got = _u32(hb.hand_potential_counts(torch.from_numpy(sit).cuda(), "treys"))
    workers = max(1, min(32, len(os.sched_getaffinity(0))))
    chunks = [np.ascontiguousarray(sit[i::workers]) for i in range(workers)]
    with mp.get_context("fork").Pool(workers) as pool:
        parts = pool.map(_oracle_chunk, [(c.tobytes(), len(c), "treys") for c in chunks])
    want = np.zeros((1024, 16), dtype=np.uint32)
    for i, p in enumerate(parts):
        want[i::workers] = np.frombuffer(p, dtype=np.uint32).reshape(-1, 16)
    assert np.array_equal(got, want)


def test_host_entry_chunked_pipeline_matches_device_entry(hb):
    """The *_host entry points stream large batches in chunks (16,384 rows for HP, 2^20 for HS), through the library's pinned
    staging for pageable caller buffers and directly for page-locked ones: a batch spanning several chunks with a ragged
    tail must equal the device entry row for row, in both modes and on both paths."""
    n = 2 * 32768 + 4097
    sit = hb.random_situations(n, seed=404)
    d_sit = torch.from_numpy(sit).cuda()
    for mode in ("treys", "reference"):
        dev = _u32(hb.hand_potential_counts(d_sit, mode))
        host = hb.hand_potential_counts_host(sit, mode)
        assert host.dtype == np.uint32 and np.array_equal(host, dev), mode
        pin_in = torch.empty((n, 8), dtype=torch.uint8, pin_memory=True)
        pin_in.numpy()[:] = sit
        pin_out = torch.zeros((n, 16), dtype=torch.int32, pin_memory=True)
        got = hb.hand_potential_counts_host(pin_in.numpy(), mode, out=pin_out.numpy().view(np.uint32))
        assert np.array_equal(got, dev), mode
        keep = pin_in.numpy()[n - 3].copy()            # a malformed row far into the batch is reported on this path too
        pin_in.numpy()[n - 3, 1] = pin_in.numpy()[n - 3, 0]   # (the kernels raise a device flag; no host-side scan otherwise)
        with pytest.raises(hb.HoldemError):
            hb.hand_potential_counts_host(pin_in.numpy(), mode, out=pin_out.numpy().view(np.uint32))
        assert bool((pin_out[n - 3] == -1).all())
        pin_in.numpy()[n - 3] = keep
    big = hb.random_situations((1 << 20) + 12345, seed=405, n_board=5)
    assert np.array_equal(hb.hand_strength_counts_host(big, "treys"),
                          _u32(hb.hand_strength_counts(torch.from_numpy(big).cuda(), "treys")))


def test_concurrent_streams_and_threads(hb, oracle):
    """Device entry points are stream-ordered and thread-safe: four host threads, each on its own CUDA stream."""
    import threading
    sit = hb.random_situations(256, seed=406)
    want = oracle.hp_batch(sit[:32], "treys")
    d_sit = torch.from_numpy(sit).cuda()
    ref = hb.hand_potential_counts(d_sit, "treys").clone()
    torch.cuda.synchronize()
    results, errors = [None] * 4, []

    def worker(i):
        try:
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                outs = [hb.hand_potential_counts(d_sit, "treys") for _ in range(5)]
                outs.append(hb.hand_potential_counts(d_sit[:64].contiguous(), "reference"))
            st.synchronize()
            results[i] = outs
        except Exception as e:  # pragma: no cover
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(4)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors
    ref64 = hb.hand_potential_counts(d_sit[:64].contiguous(), "reference", direct=True)
    for outs in results:
        assert all(torch.equal(o, ref) for o in outs[:5]) and torch.equal(outs[5], ref64)
    assert np.array_equal(_u32(ref[:32]), want)


def _texture_situations(hb):
    """Flops chosen by suit texture: monotone / two-tone / rainbow boards, hole cards inside and outside the board's
    suits, paired and tripled boards, wheel and broadway draws -- the cases where the rank-multiset kernel's flush
    corrections differ."""
    cases = [([14, 1], [31, 32, 33]), ([0, 1], [2, 3, 4]), ([0, 13], [5, 7, 9]), ([12, 11], [0, 2, 4]),
             ([26, 27], [0, 2, 17]), ([12, 25], [38, 51, 0]), ([5, 18], [31, 44, 6]), ([3, 4], [5, 6, 20]),
             ([12, 11], [10, 9, 8]), ([13, 26], [0, 1, 15]), ([0, 12], [13, 26, 39]), ([1, 14], [0, 13, 26]),
             ([12, 3], [0, 1, 2]), ([25, 38], [0, 1, 2]), ([12, 51], [0, 13, 3]), ([8, 9], [10, 11, 25]),
             ([21, 47], [8, 34, 9]), ([7, 20], [33, 46, 6]), ([39, 40], [41, 42, 0]), ([50, 51], [49, 48, 47]),
             ([0, 1], [5, 18, 31]), ([44, 2], [5, 18, 31]), ([12, 25], [38, 11, 24]), ([0, 14], [28, 42, 4])]
    return hb.pack_situations([c[0] for c in cases], [c[1] for c in cases])


def test_hp_flop_variants_agree_and_match_oracle(hb, oracle):
    """The flop kernels (rank-multiset v4 = production, 5 = without its closed-form add parts, 7 = with the step-by-step undo half
    instead of the bit-parallel one, 4-set walk v3, warp-per-pair v2) and the plain enumeration return identical rows; texture
    cases and 96 random deals are also checked against the CPU oracle."""
    sit = np.concatenate([_texture_situations(hb), hb.random_situations(96, seed=811)])
    want = oracle.hp_batch(sit, "treys")
    d_sit = torch.from_numpy(sit).cuda()
    for variant in (4, 5, 7, 3, 2):
        got = _u32(hb.hand_potential_counts(d_sit, "treys", flop_variant=variant))
        bad = np.nonzero((got != want).any(axis=1))[0]
        assert bad.size == 0, (variant, bad[:5], sit[bad[:5]], got[bad[:5]], want[bad[:5]])
    assert np.array_equal(_u32(hb.hand_potential_counts(d_sit, "treys", direct=True)), want)
    big = hb.random_situations(16384, seed=812)
    d_big = torch.from_numpy(big).cuda()
    v4 = hb.hand_potential_counts(d_big, "treys", flop_variant=4)
    v3 = hb.hand_potential_counts(d_big, "treys", flop_variant=3)
    assert torch.equal(v4, v3)
    assert torch.equal(v4, hb.hand_potential_counts(d_big, "treys", flop_variant=7))
    with pytest.raises(hb.HoldemError):
        hb.hand_potential_counts(d_sit, "treys", flop_variant=8)


def test_hp_turn_variants_agree_and_match_oracle(hb, oracle):
    """Turn rows: the rank-multiset kernel (production), the 3-set walk and the plain enumeration return identical rows;
    suit-texture cases (four / three to a flush on board, two suits twice, hole cards in and out of the suit, paired and
    tripled boards) and random deals are checked against the CPU oracle."""
    cases = [([0, 1], [2, 3, 4, 5]), ([13, 26], [2, 3, 4, 5]), ([12, 11], [2, 3, 4, 18]), ([25, 38], [2, 3, 4, 18]),
             ([0, 13], [2, 3, 15, 16]), ([12, 25], [38, 51, 0, 13]), ([5, 18], [31, 44, 6, 19]), ([7, 8], [9, 10, 24, 37]),
             ([39, 40], [41, 42, 43, 0]), ([39, 1], [41, 42, 43, 44]), ([50, 51], [49, 48, 47, 46]), ([14, 1], [31, 32, 33, 20]),
             ([0, 12], [13, 26, 39, 1]), ([21, 47], [8, 34, 9, 35]), ([0, 1], [5, 18, 31, 44]), ([44, 2], [5, 18, 31, 6]),
             ([12, 25], [38, 11, 24, 37])]
    sit = np.concatenate([hb.pack_situations([c[0] for c in cases], [c[1] for c in cases]),
                          hb.random_situations(400, seed=821, n_board=4)])
    want = oracle.hp_batch(sit, "treys")
    d_sit = torch.from_numpy(sit).cuda()
    for variant in (4, 3):
        got = _u32(hb.hand_potential_counts(d_sit, "treys", turn_variant=variant))
        bad = np.nonzero((got != want).any(axis=1))[0]
        assert bad.size == 0, (variant, bad[:5], sit[bad[:5]], got[bad[:5]], want[bad[:5]])
    assert np.array_equal(_u32(hb.hand_potential_counts(d_sit, "treys")), want)
    big = hb.random_situations(65536, seed=822, n_board=4)
    d_big = torch.from_numpy(big).cuda()
    assert torch.equal(hb.hand_potential_counts(d_big, "treys", turn_variant=4),
                       hb.hand_potential_counts(d_big, "treys", turn_variant=3))


def test_hp_batch_peers_replicates_rows(hb, oracle):
    """holdem_hp_batch_peers: the kernels store every result row into the peer buffers as well (over NVLink when the peers are
    other GPUs' symmetric memory; here two more buffers on the same device stand in).  Mixed streets, both modes, with a row
    offset into the peer buffers as the sharded layer uses it."""
    import ctypes
    from holdem_b200 import _lib
    from holdem_b200.api import _enter
    parts = [hb.random_situations(300, seed=901), hb.random_situations(200, seed=902, n_board=4),
             hb.random_situations(50, seed=903, n_board=5)]
    sit = np.concatenate(parts)
    sit = sit[np.random.default_rng(9).permutation(len(sit))]
    sit[7, 1] = sit[7, 0]                                  # one malformed row: must be marked in every buffer
    n, off = len(sit), 11
    d_sit = torch.from_numpy(sit).cuda()
    for mode in ("treys", "reference"):
        sub = d_sit if mode == "treys" else d_sit[:96].contiguous()
        m = sub.shape[0]
        want = hb.hand_potential_counts(sub, mode)
        mine = torch.zeros((m, 16), dtype=torch.int32, device="cuda")
        peers = [torch.full((m + 2 * off, 16), 7, dtype=torch.int32, device="cuda") for _ in range(2)]
        with _enter(sub) as (lib, st):
            arr = (ctypes.c_void_p * 2)(*[p.data_ptr() + off * 64 for p in peers])
            _lib.check(lib.holdem_hp_batch_peers(sub.data_ptr(), m, hb.mode_id(mode), mine.data_ptr(), arr, 2, st))
        torch.cuda.synchronize()
        # the torch op over the same entry point
        mine2 = torch.zeros_like(mine)
        peers2 = [torch.full((m + 2 * off, 16), 7, dtype=torch.int32, device="cuda") for _ in range(2)]
        torch.ops.holdem.hp_peers(sub, hb.mode_id(mode), mine2, [p.data_ptr() + off * 64 for p in peers2])
        assert torch.equal(mine2, want) and all(torch.equal(a, b) for a, b in zip(peers, peers2)), mode
        assert torch.equal(mine, want), mode
        for p in peers:
            assert torch.equal(p[off:off + m], want), mode
            assert bool((p[:off] == 7).all()) and bool((p[off + m:] == 7).all()), mode   # nothing outside the block is touched
    assert np.array_equal(_u32(want[8:32]), oracle.hp_batch(sit[8:32], "reference"))       # (row 7 is the malformed one)
    assert bool((want[7] == -1).all())


def test_hs_flush_board_textures_match_oracle(hb, oracle):
    """Treys-mode hand strength comes from rank pairs plus a correction for holdings that make a flush with the board: boards with
    three, four and five cards of a suit (hole cards inside and outside it, paired boards) against the oracle, row by row."""
    cases = [([14, 1], [31, 32, 33]), ([0, 1], [2, 3, 4]), ([13, 26], [2, 3, 4]), ([0, 1], [2, 3, 4, 5]), ([13, 26], [2, 3, 4, 5]),
             ([12, 11], [2, 3, 4, 18]), ([25, 38], [2, 3, 4, 18]), ([39, 1], [41, 42, 43, 44]), ([0, 1], [2, 3, 4, 5, 6]),
             ([13, 26], [2, 3, 4, 5, 7]), ([12, 25], [0, 2, 4, 6, 8]), ([11, 13], [0, 2, 4, 6, 21]), ([0, 13], [26, 27, 28, 29, 30]),
             ([20, 39], [40, 13, 35, 30, 28]), ([5, 18], [31, 44, 6, 19, 32]), ([50, 51], [49, 48, 47, 46, 45]),
             ([12, 10], [51, 50, 49, 48, 47]), ([0, 14], [39, 40, 41, 42, 1])]
    for nb in (3, 4, 5):
        sub = [c for c in cases if len(c[1]) == nb]
        sit = np.concatenate([hb.pack_situations([c[0] for c in sub], [c[1] for c in sub]),
                              hb.random_situations(2000, seed=930 + nb, n_board=nb)])
        got = _u32(hb.hand_strength_counts(torch.from_numpy(sit).cuda(), "treys"))
        want = oracle.hs_batch(sit, "treys")
        bad = np.nonzero((got != want).any(axis=1))[0]
        assert bad.size == 0, (nb, sit[bad[:5]], got[bad[:5]], want[bad[:5]])


def test_full_size_configs_reproduce_committed_hashes(hb):
    """BASELINE.json configs[2], configs[3] (22 seats x 100,000 hands, every street) and configs[4] (1,000,000 flop situations) at
    FULL size: the 64-bit content hash of every result buffer equals the one committed in tests/golden/gpu_hashes.json (written
    by scripts/make_gpu_hashes.py from a run in which the production kernels equalled the independent kernels on all rows and the
    oracle on samples); and the sum of counts, by contrast, is the same structural constant for any well-formed rows."""
    from holdem_b200 import table
    g = load_golden("gpu_hashes.json")

    def hx(t):
        return f"{hb.rows_hash(t.contiguous()):016x}"

    sit = torch.from_numpy(hb.random_situations(65536, seed=20251019)).cuda()
    assert hx(hb.hand_potential_counts(sit, "treys")) == g["configs2_first_65536_rows_treys"]
    assert hx(hb.hand_potential_counts(sit, "reference")) == g["configs2_first_65536_rows_reference"]
    sit = torch.from_numpy(hb.random_situations(1_000_000, seed=20251021)).cuda()
    rows = hb.hand_potential_counts(sit, "treys")
    assert hx(rows) == g["configs4_1m_flop_rows_treys"]
    assert int(rows[:, :15].to(torch.int64).sum()) == 1_000_000 * 1_072_352
    del rows, sit
    hole, board = table.deal_tables(100_000, seed=20251020, seats=22)
    d_hole, d_board = torch.from_numpy(hole).cuda(), torch.from_numpy(board).cuda()
    for nb, name in ((3, "flop_hs_hp"), (4, "turn_hs_hp"), (5, "river_hs")):
        assert hx(hb.table_hand_potential_counts(d_hole, d_board, nb, "treys").reshape(-1, 16)) == g[f"configs3_22seat_100k_{name}"], name
