"""GPU parity: batched 5/6/7-card Treys evaluator and literal category ranker vs the CPU oracle, through the C ABI."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _random_hands(n, n_cards, seed):
    rng = np.random.default_rng(seed)
    cards = np.argsort(rng.random((n, 52)), axis=1)[:, :n_cards].astype(np.uint8)
    rows = np.full((n, 8), 0xFF, dtype=np.uint8)
    rows[:, :n_cards] = cards
    return rows


@pytest.mark.parametrize("n_cards", [5, 6, 7])
@pytest.mark.parametrize("n", [1, 1000, 400_000])   # small -> L1/L2-table kernel, large -> shared-memory-table kernel
def test_eval_matches_oracle(hb, oracle, n_cards, n):
    rows = _random_hands(n, n_cards, seed=100 + n_cards)
    got = hb.evaluate(torch.from_numpy(rows).cuda(), n_cards).cpu().numpy().astype(np.uint16)
    assert np.array_equal(got, oracle.treys_evaluate_batch(rows, n_cards))


def test_eval_all_five_card_hands(hb, oracle):
    import itertools
    rows = np.array([list(c) + [255, 255, 255] for c in itertools.combinations(range(52), 5)], dtype=np.uint8)
    got = hb.evaluate(torch.from_numpy(rows).cuda(), 5).cpu().numpy().astype(np.uint16)
    assert np.array_equal(got, oracle.treys_evaluate_batch(rows, 5))
    assert len(np.unique(got)) == 7462


def test_eval_edge_cases(hb):
    rows = np.array([[0, 1, 2, 3, 4, 5, 6, 255],          # ok
                     [0, 0, 2, 3, 4, 5, 6, 255],          # duplicate
                     [0, 1, 2, 3, 4, 5, 52, 255],         # unknown card
                     [0, 1, 2, 3, 4, 5, 255, 255]],       # too few cards for n=7
                    dtype=np.uint8)
    got = hb.evaluate(torch.from_numpy(rows).cuda(), 7).cpu().tolist()
    assert got[0] > 0 and got[1:] == [0, 0, 0]
    assert hb.evaluate(torch.empty((0, 8), dtype=torch.uint8, device="cuda"), 7).numel() == 0
    with pytest.raises(KeyError):
        hb.evaluate(torch.from_numpy(rows).cuda(), 4)
    with pytest.raises(RuntimeError):
        hb.evaluate(torch.from_numpy(rows), 7)            # CPU tensor: no fallback
    with pytest.raises(hb.HoldemError):
        hb.evaluate_host(rows, 7)                          # malformed rows are an error at the host boundary
    assert hb.evaluate_host(rows[:1], 7).tolist() == [got[0]]


def test_eval_host_entry(hb, oracle):
    rows = _random_hands(300_000, 7, seed=3)
    assert np.array_equal(hb.evaluate_host(rows, 7), oracle.treys_evaluate_batch(rows, 7))


def test_exhaustive_seven_cards(hb):
    """BASELINE.json configs[1]: all 133,784,560 hands, published class frequencies and checksums."""
    hist, sums = hb.exhaustive7()
    # the whole rank histogram equals the CPU oracle's, bin by bin (hash committed from tests/test_oracle_treys.py)
    assert f"{hb.rows_hash(hist.view(torch.int32).reshape(-1, 2)):016x}" == load_golden("gpu_hashes.json")["exhaustive7_rank_histogram"]
    hist = hist.cpu().numpy()
    assert int(hist.sum()) == 133784560
    bounds = [0, 10, 166, 322, 1599, 1609, 2467, 3325, 6185, 7462]
    per_class = [int(hist[bounds[i] + 1:bounds[i + 1] + 1].sum()) for i in range(9)]
    assert per_class == [41584, 224848, 3473184, 4047644, 6180020, 6461620, 31433400, 58627800, 23294460]
    assert int(hist[0]) == 0 and int(hist[1]) == 4324 and int(hist[10]) == 4140 and int(hist[11]) == 4052
    assert int(hist[1600]) == 747980 and int(hist[7462]) == 0 and int((hist > 0).sum()) == 4824
    assert sums.cpu().tolist() == [547965983972, 2665153084455708]


def test_category_golden_and_random(hb, oracle):
    g = load_golden("literal_category.json")["cases"]
    for n_cards in range(5, 10):
        cases = [c for c in g if len(c[0]) == n_cards]
        rows = np.full((len(cases), 16), 0xFF, dtype=np.uint8)
        for i, (cards, _) in enumerate(cases):
            rows[i, :n_cards] = cards
        got = hb.category(torch.from_numpy(rows).cuda(), n_cards).cpu().tolist()
        assert got == [v for _, v in cases], n_cards
    rng = np.random.default_rng(11)
    for n_cards in range(2, 10):
        rows = rng.integers(0, 52, size=(200_000, 9), dtype=np.uint8)   # duplicates on purpose
        rows[::2, :] = rng.integers(0, 4, size=(100_000, 9), dtype=np.uint8) * 13 + \
            rng.integers(0, 5, size=(100_000, 9), dtype=np.uint8)        # few ranks: quads/full houses/5-of-a-value
        got = hb.category(torch.from_numpy(rows).cuda(), n_cards).cpu().numpy()
        assert np.array_equal(got, oracle.lit_category_batch(rows, n_cards)), n_cards
    bad = np.array([[0, 1, 2, 3, 60, 0, 0, 0, 0]], dtype=np.uint8)
    assert hb.category(torch.from_numpy(bad).cuda(), 5).cpu().tolist() == [0xFF]
    assert np.array_equal(hb.category_host(rows[:1000], 7), oracle.lit_category_batch(rows[:1000], 7))


@pytest.mark.parametrize("n_cards", [5, 6, 7])
def test_eval_large_batch_fast_and_strict_paths(hb, oracle, n_cards):
    """Large batches take the fast kernel (suit counters + deferred flush queue): every rank equals the oracle, cards
    > 51 are flagged in both paths, duplicate cards only by strict=True / the host entry (like Treys, the fast path
    does not look for them)."""
    n = 700_003                                      # not a multiple of anything: exercises the ragged tail
    rows = _random_hands(n, n_cards, seed=500 + n_cards)
    rng = np.random.default_rng(1)
    flushy = rng.integers(0, n, 20000)               # force plenty of flushes / straight flushes
    suit = rng.integers(0, 4, 20000)
    for i, s in zip(flushy, suit):
        rows[i, :n_cards] = 13 * s + rng.permutation(13)[:n_cards]
    want = oracle.treys_evaluate_batch(rows, n_cards)
    bad_card = rng.integers(0, n, 500)
    dup = np.setdiff1d(rng.integers(0, n, 500), bad_card)
    rows[bad_card, rng.integers(0, n_cards, 500)] = rng.integers(52, 256, 500).astype(np.uint8)
    rows[dup, 0] = rows[dup, 1]
    d = torch.from_numpy(rows).cuda()
    fast = hb.evaluate(d, n_cards).cpu().numpy().astype(np.uint16)
    strict = hb.evaluate(d, n_cards, strict=True).cpu().numpy().astype(np.uint16)
    good = np.ones(n, dtype=bool)
    good[bad_card] = False
    good[dup] = False
    assert np.array_equal(fast[good], want[good]) and np.array_equal(strict[good], want[good])
    assert (fast[bad_card] == 0).all() and (strict[bad_card] == 0).all() and (strict[dup] == 0).all()


@pytest.mark.parametrize("n_cards", [5, 7])
def test_eval_fast_path_deferral_extremes(hb, oracle, n_cards):
    """The fast kernel defers flush-pattern hands through per-lane windows and a per-warp stack.  Dense runs of flushes
    overflow the stack (each lane then finishes its own hands), sparse ones are carried from window to window, and
    a batch of nothing but deferred hands must still come out right: compared with the strict kernel on every row and
    with the oracle on a sample."""
    n = 1_000_003
    rng = np.random.default_rng(77 + n_cards)
    rows = _random_hands(n, n_cards, seed=900 + n_cards)
    suited = (13 * rng.integers(0, 4, n)[:, None] + np.argsort(rng.random((n, 13)), axis=1)[:, :n_cards]).astype(np.uint8)
    dense = np.zeros(n, dtype=bool)
    dense[100_000:400_000] = True                                  # every hand deferred
    dense[400_000:600_000] = rng.random(200_000) < 0.5             # half of them
    dense[600_000:700_000] = (np.arange(100_000) % 151_552) < 64   # short bursts
    rows[dense, :n_cards] = suited[dense]
    rows[700_000:700_512, 0] = 200                                 # a run of out-of-range bytes
    d = torch.from_numpy(rows).cuda()
    fast = hb.evaluate(d, n_cards).cpu().numpy().astype(np.uint16)
    strict = hb.evaluate(d, n_cards, strict=True).cpu().numpy().astype(np.uint16)
    assert np.array_equal(fast, strict)
    assert (fast[700_000:700_512] == 0).all() and (fast[:700_000] != 0).all()
    sample = np.r_[0:2000, 100_000:102_000, 400_000:402_000, n - 2000:n]
    assert np.array_equal(fast[sample], oracle.treys_evaluate_batch(rows[sample], n_cards))
    # all-deferred batch, just above the size at which the fast kernel is chosen
    m = 148 * 2048 + 5
    d2 = torch.from_numpy(np.ascontiguousarray(rows[100_000:100_000 + m])).cuda()
    assert torch.equal(hb.evaluate(d2, n_cards), hb.evaluate(d2, n_cards, strict=True))


def test_showdown_whole_table(hb, oracle):
    """all 22 seats of a hand in one evaluator launch; winners under the correct rule and under the
    reference's max() rule (GameBoard.py:210)."""
    from holdem_b200 import table
    hole, board = table.deal_tables(2000, seed=20251020, seats=22)
    ranks = table.showdown_ranks(torch.from_numpy(hole).cuda(), torch.from_numpy(board).cuda())
    rows = np.full((2000 * 22, 8), 255, dtype=np.uint8)
    rows[:, 0:2] = hole.reshape(-1, 2)
    rows[:, 2:7] = np.repeat(board, 22, axis=0)
    want = oracle.treys_evaluate_batch(rows, 7).reshape(2000, 22)
    assert np.array_equal(ranks.cpu().numpy().astype(np.uint16), want)
    best = table.winners(ranks, "best").cpu().numpy()
    ref = table.winners(ranks, "reference").cpu().numpy()
    assert np.array_equal(best, want == want.min(axis=1, keepdims=True))
    assert np.array_equal(ref, want == want.max(axis=1, keepdims=True))


def test_fused_showdown_kernel_equals_evaluator_path(hb, oracle):
    """The one-kernel showdown (a warp per hand, a lane per seat) returns the evaluator's ranks and the winner masks of
    table.winners under both rules, with and without folded seats, through both bindings; ranks pinned by the oracle."""
    from holdem_b200 import table
    for seats, hands in ((22, 3000), (9, 1000), (2, 257), (1, 5)):
        hole, board = table.deal_tables(hands, seed=100 + seats, seats=max(seats, 2))
        hole = hole[:, :seats].copy()
        h, b = torch.from_numpy(hole).cuda(), torch.from_numpy(board).cuda()
        ranks, best, ref = table.showdown(h, b)
        want = table.showdown_ranks(h, b)
        assert torch.equal(ranks, want)
        assert torch.equal(best, table.winners(want, "best")) and torch.equal(ref, table.winners(want, "reference"))
        rows = np.full((hands * seats, 8), 0xFF, dtype=np.uint8)
        rows[:, 0:2] = hole.reshape(-1, 2)
        rows[:, 2:7] = np.repeat(board, seats, axis=0)
        assert np.array_equal(ranks.cpu().numpy().astype(np.uint16).ravel(), oracle.treys_evaluate_batch(rows, 7))
        act = torch.from_numpy(np.random.default_rng(seats).random((hands, seats)) < 0.6).cuda()
        r2, best2, ref2 = table.showdown(h, b, act)
        assert torch.equal(r2, want)
        assert torch.equal(best2, table.winners(want, "best", act)) and torch.equal(ref2, table.winners(want, "reference", act))
        rc, mc = hb.api.table_showdown(h, b, act, binding="ctypes")
        ro, mo = hb.api.table_showdown(h, b, act)
        assert torch.equal(rc, ro) and torch.equal(mc, mo)
    # a malformed seat (duplicate card) ranks 0 and never wins; an empty batch is a no-op
    hole, board = table.deal_tables(4, seed=1, seats=4)
    hole[2, 1] = [board[2, 0], hole[2, 1, 1]]
    ranks, best, ref = table.showdown(torch.from_numpy(hole).cuda(), torch.from_numpy(board).cuda())
    assert int(ranks[2, 1]) == 0 and not bool(best[2, 1]) and not bool(ref[2, 1]) and bool(best[2].any())
    e = table.showdown(torch.empty((0, 22, 2), dtype=torch.uint8).cuda(), torch.empty((0, 5), dtype=torch.uint8).cuda())
    assert e[0].shape == (0, 22) and e[1].shape == (0, 22)
    with pytest.raises((RuntimeError, ValueError)):
        hb.api.table_showdown(torch.zeros((1, 23, 2), dtype=torch.uint8).cuda(), torch.zeros((1, 5), dtype=torch.uint8).cuda(),
                              binding="ctypes")
