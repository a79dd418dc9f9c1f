"""Extensions around the hot path that the reference lists as future work (README.md:106-114): hand
strength against several opponents, pre-flop hand strength over sampled flops, and Monte-Carlo all-in equity against up
to 21 opponents.  None of this exists in the reference (parity unpinned); everything is built from the same kernels
(hs_kernel / the batched evaluator) and checked against exact enumeration where one exists (tests/test_gpu_extensions.py).
The sampling itself is torch on the device (uniform without replacement = top-k of uniform keys); no CPU fallback.
"""
import numpy as np
import torch

from .api import evaluate, hand_strength_counts, hs_from_counts

__all__ = ["hand_strength_vs_n", "preflop_hand_strength", "monte_carlo_equity"]


def hand_strength_vs_n(hs, n_opponents: int):
    """HS^n: probability of being ahead of n independent random holdings (the usual multi-opponent approximation;
    Effective-hand-strength article linked at HandCalculations.py:14).  hs: float or tensor."""
    if n_opponents < 1:
        raise ValueError("n_opponents must be >= 1")
    return hs ** n_opponents


def _check_cards(cards, lo, hi, name):
    cards = [int(c) for c in cards]
    if not lo <= len(cards) <= hi or any(c < 0 or c > 51 for c in cards) or len(set(cards)) != len(cards):
        raise ValueError(f"{name}: expected {lo}..{hi} distinct cards in [0,52)")
    return cards


def preflop_hand_strength(hole, n_flops: int = 4096, seed: int = 0, mode="treys", device=None):
    """Mean flop HandStrength of the two hole cards over `n_flops` flops sampled uniformly from the other 50 cards.
    Returns (mean HS as a Python float, per-flop HS tensor [n_flops] float64 on the device)."""
    hole = _check_cards(hole, 2, 2, "hole")
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    rest = torch.tensor([c for c in range(52) if c not in hole], dtype=torch.uint8, device=dev)
    pick = torch.rand((n_flops, 50), device=dev, generator=g).topk(3, dim=1).indices
    rows = torch.full((n_flops, 8), 0xFF, dtype=torch.uint8, device=dev)
    rows[:, 0] = hole[0]
    rows[:, 1] = hole[1]
    rows[:, 2:5] = rest[pick]
    hs = hs_from_counts(hand_strength_counts(rows, mode))
    return float(hs.mean()), hs


def monte_carlo_equity(hole, board=(), n_opponents: int = 1, n_samples: int = 65536, seed: int = 0, device=None):
    """All-in equity of `hole` on `board` (0, 3, 4 or 5 cards) against n_opponents (1..21) random holdings: the rest of
    the board and every opponent's two cards are dealt uniformly from the unseen cards, all 7-card hands are ranked by
    the batched evaluator kernel (Treys rank, lower wins).
    -> dict(win, tie, lose: fractions of samples; equity: mean pot share, ties split evenly; samples)."""
    hole = _check_cards(hole, 2, 2, "hole")
    board = _check_cards(board, 0, 5, "board")
    if len(board) in (1, 2) or set(hole) & set(board):
        raise ValueError("board must hold 0, 3, 4 or 5 cards distinct from the hole cards")
    if not 1 <= n_opponents <= 21:
        raise ValueError("n_opponents must be 1..21 (GameBoard.MAX_PLAYERS = 22)")
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    known = hole + board
    rest = torch.tensor([c for c in range(52) if c not in known], dtype=torch.uint8, device=dev)
    n_board_draw = 5 - len(board)
    need = n_board_draw + 2 * n_opponents
    pick = rest[torch.rand((n_samples, rest.numel()), device=dev, generator=g).topk(need, dim=1).indices]
    full_board = torch.empty((n_samples, 5), dtype=torch.uint8, device=dev)
    if board:
        full_board[:, :len(board)] = torch.tensor(board, dtype=torch.uint8, device=dev)
    full_board[:, len(board):] = pick[:, :n_board_draw]
    seats = n_opponents + 1
    rows = torch.full((n_samples, seats, 8), 0xFF, dtype=torch.uint8, device=dev)
    rows[:, 0, 0] = hole[0]
    rows[:, 0, 1] = hole[1]
    rows[:, 1:, 0:2] = pick[:, n_board_draw:].reshape(n_samples, n_opponents, 2)
    rows[:, :, 2:7] = full_board[:, None, :]
    ranks = evaluate(rows.reshape(n_samples * seats, 8), 7).reshape(n_samples, seats).to(torch.int32)
    ours, best_opp = ranks[:, 0], ranks[:, 1:].min(dim=1).values
    win = ours < best_opp
    tie = ours == best_opp
    n_tied = (ranks[:, 1:] == ours[:, None]).sum(dim=1) + 1          # seats sharing the pot when we tie for best
    share = torch.where(win, torch.ones_like(ours, dtype=torch.float64),
                        torch.where(tie, 1.0 / n_tied.to(torch.float64), torch.zeros_like(ours, dtype=torch.float64)))
    return {"win": float(win.double().mean()), "tie": float(tie.double().mean()),
            "lose": float((~win & ~tie).double().mean()), "equity": float(share.mean()), "samples": int(n_samples)}
