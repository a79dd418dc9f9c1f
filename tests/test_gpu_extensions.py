"""SURVEY.md 8f4 extensions (multi-opponent HS, pre-flop HS over sampled flops, Monte-Carlo equity): built from the
path's kernels and checked against exact enumeration where one exists.  Parity with the reference is unpinned -- the
reference lists these as future work (README.md:106-114)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_monte_carlo_equity_matches_exact_hand_strength_on_the_river(hb, oracle):
    """One opponent, complete board: the equity is exactly (ahead + tied/2) / 990 of the enumeration."""
    hole, board = [20, 39], [40, 13, 35, 30, 28]
    a, t, b = oracle.hs_counts(hole, board, "treys")
    exact = (a + t / 2) / (a + t + b)
    mc = hb.monte_carlo_equity(hole, board, n_opponents=1, n_samples=400000, seed=3)
    assert abs(mc["equity"] - exact) < 0.004 and abs(mc["win"] - a / 990) < 0.004 and abs(mc["tie"] - t / 990) < 0.004
    assert abs(mc["win"] + mc["tie"] + mc["lose"] - 1.0) < 1e-12
    again = hb.monte_carlo_equity(hole, board, n_opponents=1, n_samples=400000, seed=3)
    assert again == mc                                                     # seeded: reproducible


def test_monte_carlo_equity_flop_one_opponent_matches_hp_counts(hb):
    """One opponent on the flop: P(we end ahead) from the full HS+HP enumeration = sum of the 'ahead' column / 1,070,190."""
    hole, board = [14, 1], [31, 32, 33]
    rows = hb.hand_potential_counts(torch.from_numpy(hb.pack_situations([hole], [board])).cuda(), "treys")[0].cpu().numpy()
    hp = rows[3:12].reshape(3, 3).astype(np.float64)
    total = hp.sum()
    assert total == 1081 * 990
    exact_win, exact_tie = hp[:, 0].sum() / total, hp[:, 1].sum() / total
    mc = hb.monte_carlo_equity(hole, board, n_opponents=1, n_samples=400000, seed=5)
    assert abs(mc["win"] - exact_win) < 0.004 and abs(mc["tie"] - exact_tie) < 0.004


def test_more_opponents_lower_equity_and_hs_power(hb):
    hole = [12, 25]                                                        # Ac Ad
    eq = [hb.monte_carlo_equity(hole, (), n_opponents=n, n_samples=100000, seed=7)["equity"] for n in (1, 3, 9, 21)]
    assert eq[0] > eq[1] > eq[2] > eq[3] and 0.84 < eq[0] < 0.86            # aces vs one random hand: ~85.2 %
    assert hb.hand_strength_vs_n(0.8, 3) == pytest.approx(0.512)
    t = hb.hand_strength_vs_n(torch.tensor([0.5, 1.0], dtype=torch.float64), 2)
    assert t.tolist() == [0.25, 1.0]
    with pytest.raises(ValueError):
        hb.monte_carlo_equity(hole, [0, 1], n_opponents=1)
    with pytest.raises(ValueError):
        hb.monte_carlo_equity(hole, (), n_opponents=22)


def test_preflop_hand_strength_orders_hands(hb):
    aces, _ = hb.preflop_hand_strength([12, 25], n_flops=4096, seed=1)
    seven_deuce, per_flop = hb.preflop_hand_strength([5, 13], n_flops=4096, seed=1)   # 7c 2d
    assert aces > 0.8 > 0.5 > seven_deuce and per_flop.shape == (4096,)
    assert bool(((per_flop >= 0) & (per_flop <= 1)).all())
