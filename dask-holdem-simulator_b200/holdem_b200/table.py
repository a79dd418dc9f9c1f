"""Whole-table helpers around the evaluator kernel: every seat of a hand shares the board.

* deal_tables(): BASELINE configs[3] generator -- one shuffle per hand, seat i holds cards 2i, 2i+1, board = cards 44..48.
* showdown_ranks(): Treys rank of every seat's 7 cards in one evaluator launch (what GameBoard.determine_winner,
  GameBoard.py:204-229, obtains seat by seat through evaluate_best_hand).
* showdown(): ranks AND both winner masks of every hand from one fused kernel (a warp per hand, a lane per seat).
* winners(): boolean winner mask.  rule="best": lowest Treys rank wins (correct poker); rule="reference": the seats the
  reference's determine_winner picks -- it takes max() of the Treys values (GameBoard.py:210), i.e. the WORST hands.
"""
import numpy as np
import torch

from .api import evaluate, table_showdown


def deal_tables(n_hands: int, seed: int = 20251020, seats: int = 22):
    """-> (hole uint8 [n_hands, seats, 2], board uint8 [n_hands, 5]) as numpy arrays."""
    if not 2 <= seats <= 22:
        raise ValueError("seats must be 2..22 (GameBoard.MAX_PLAYERS = 22)")
    rng = np.random.default_rng(seed)
    hole = np.empty((n_hands, seats, 2), dtype=np.uint8)
    board = np.empty((n_hands, 5), dtype=np.uint8)
    step = 20000
    for s in range(0, n_hands, step):
        m = min(step, n_hands - s)
        deck = np.argsort(rng.random((m, 52)), axis=1).astype(np.uint8)
        hole[s:s + m] = deck[:, :2 * seats].reshape(m, seats, 2)
        board[s:s + m] = deck[:, 44:49]
    return hole, board


def situations_for_street(hole: np.ndarray, board: np.ndarray, n_board: int) -> np.ndarray:
    """Situation rows uint8 [n_hands * seats, 8] for every seat with the first n_board (3/4/5) board cards."""
    n, seats = hole.shape[0], hole.shape[1]
    rows = np.full((n, seats, 8), 0xFF, dtype=np.uint8)
    rows[:, :, 0:2] = hole
    rows[:, :, 2:2 + n_board] = board[:, None, :n_board]
    return rows.reshape(n * seats, 8)


def showdown_ranks(hole: torch.Tensor, board: torch.Tensor) -> torch.Tensor:
    """hole uint8 [N,S,2], board uint8 [N,5] (CUDA) -> int16 [N,S] Treys ranks (1..7462, lower is better)."""
    n, seats = hole.shape[0], hole.shape[1]
    rows = torch.full((n, seats, 8), 0xFF, dtype=torch.uint8, device=hole.device)
    rows[:, :, 0:2] = hole
    rows[:, :, 2:7] = board[:, None, :]
    return evaluate(rows.reshape(n * seats, 8), 7).reshape(n, seats)


def showdown(hole: torch.Tensor, board: torch.Tensor, active: torch.Tensor = None):
    """-> (ranks int16 [N,S], best bool [N,S], reference bool [N,S]) from the fused showdown kernel: `best` marks the seats that
    take (or split) the pot, `reference` the seats GameBoard.determine_winner picks (max() of the Treys values, GameBoard.py:210)."""
    ranks, masks = table_showdown(hole, board, active)
    bit = torch.arange(hole.shape[1], device=hole.device, dtype=torch.int32)
    return ranks, ((masks[:, 0:1] >> bit) & 1).bool(), ((masks[:, 1:2] >> bit) & 1).bool()


def winners(ranks: torch.Tensor, rule: str = "best", active: torch.Tensor = None) -> torch.Tensor:
    """Boolean [N,S] mask of the seats that take (or split) the pot among `active` seats (default: all)."""
    r = ranks.to(torch.int32)
    if active is None:
        active = torch.ones_like(r, dtype=torch.bool)
    if rule == "best":
        target = torch.where(active, r, torch.full_like(r, 1 << 20)).min(dim=1, keepdim=True).values
    elif rule == "reference":
        target = torch.where(active, r, torch.full_like(r, -1)).max(dim=1, keepdim=True).values
    else:
        raise ValueError("rule must be 'best' or 'reference'")
    return active & (r == target)
