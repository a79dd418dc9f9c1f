"""Card encodings met on the path and converters between them.

* path encoding (HandCalculations.py:53-54): card in [0,52), rank = card % 13 (0 = deuce), suit = card // 13
  (0 c, 1 d, 2 h, 3 s; testing/ExampleHS.py:4-5).  This is what every kernel consumes.
* rank-major encoding (utils.py:63-67, :82-85): rank * 4 + suit, used by GameBoard/Player messages.
* Treys card int (treys card.py, reached via pokerlib_to_treys.py:12-13):
  (1 << rank) << 16 | suit_bit << 12 | rank << 8 | prime[rank], suit_bit s=1 h=2 d=4 c=8.
* poker.Card objects: duck-typed through `.rank.val` ('2'..'A') and `.suit.value[1]` ('c','d','h','s'), exactly the
  two attributes pokerlib_to_treys.convert_to_treys reads.
"""
import numpy as np
import torch

RANK_CHARS = "23456789TJQKA"
SUIT_CHARS = "cdhs"
PRIMES = (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41)
_TREYS_SUIT_BIT = {"s": 1, "h": 2, "d": 4, "c": 8}
_SUIT_FROM_BIT = {1: 3, 2: 2, 4: 1, 8: 0}


def path_from_str(s: str) -> int:
    """'As' -> 51, '2c' -> 0."""
    return SUIT_CHARS.index(s[1]) * 13 + RANK_CHARS.index(s[0])


def str_from_path(c: int) -> str:
    return RANK_CHARS[c % 13] + SUIT_CHARS[c // 13]


def treys_from_str(s: str) -> int:
    """treys.Card.new: 'Kd' -> 0x08004B25."""
    r = RANK_CHARS.index(s[0])
    return ((1 << r) << 16) | (_TREYS_SUIT_BIT[s[1]] << 12) | (r << 8) | PRIMES[r]


def treys_from_path(c: int) -> int:
    return treys_from_str(str_from_path(int(c)))


def path_from_treys(t: int) -> int:
    t = int(t)
    r = (t >> 8) & 0xF
    suit_bit = (t >> 12) & 0xF
    if r > 12 or suit_bit not in _SUIT_FROM_BIT or (t >> 16) != (1 << r) or (t & 0xFF) != PRIMES[r]:
        raise ValueError(f"not a Treys card int: {t:#x}")
    return _SUIT_FROM_BIT[suit_bit] * 13 + r


def path_from_rank_major(v: int) -> int:
    """utils.card_to_int encoding (rank*4+suit) -> path encoding (suit*13+rank)."""
    r, s = divmod(int(v), 4)
    return s * 13 + r


def rank_major_from_path(c: int) -> int:
    return (int(c) % 13) * 4 + int(c) // 13


def path_from_poker_card(card) -> int:
    """poker.Card (or any object with .rank.val and .suit.value[1]) -> path encoding."""
    return path_from_str(card.rank.val + card.suit.value[1])


def pack_situations(our, board) -> np.ndarray:
    """our: int array [N,2]; board: int array [N,3|4|5] -> uint8 [N,8] rows (0xFF padded)."""
    our = np.asarray(our)
    board = np.asarray(board)
    if our.ndim == 1:
        our = our[None, :]
    if board.ndim == 1:
        board = board[None, :]
    n = our.shape[0]
    if our.shape != (n, 2) or board.shape[0] != n or not 3 <= board.shape[1] <= 5:
        raise ValueError("expected our [N,2] and board [N,3..5]")
    # validate BEFORE the uint8 cast: 256 would wrap to card 0, -1 to the 0xFF "absent" marker
    for name, a in (("our", our), ("board", board)):
        if a.size and (not np.issubdtype(a.dtype, np.integer) or a.min() < 0 or a.max() > 51):
            raise ValueError(f"{name}: cards must be integers in 0..51")
    rows = np.full((n, 8), 0xFF, dtype=np.uint8)
    rows[:, 0:2] = our
    rows[:, 2:2 + board.shape[1]] = board
    return rows


def random_situations(n: int, seed: int, n_board: int = 3) -> np.ndarray:
    """SURVEY.md section 8d deal generator: first 2+n_board cards of a uniform permutation per row."""
    rng = np.random.default_rng(seed)
    cards = np.argsort(rng.random((n, 52)), axis=1)[:, :2 + n_board].astype(np.uint8)
    rows = np.full((n, 8), 0xFF, dtype=np.uint8)
    rows[:, :2 + n_board] = cards
    return rows


# ---- tensor-in / tensor-out converters (vectorised; any device, any integer dtype) ----------------------------------------
def path_from_rank_major_t(v: torch.Tensor) -> torch.Tensor:
    """utils.card_to_int ints (rank*4+suit, utils.py:82-85) -> path encoding (suit*13+rank), elementwise."""
    return (v % 4) * 13 + torch.div(v, 4, rounding_mode="floor")


def rank_major_from_path_t(c: torch.Tensor) -> torch.Tensor:
    return (c % 13) * 4 + torch.div(c, 13, rounding_mode="floor")


_PRIMES_T = torch.tensor(PRIMES, dtype=torch.int64)
_TREYS_SUIT_BITS_T = torch.tensor([8, 4, 2, 1], dtype=torch.int64)      # path suits c d h s -> Treys suit bits


def treys_from_path_t(c: torch.Tensor) -> torch.Tensor:
    """path encoding -> Treys card ints (int64), elementwise: (1<<rank)<<16 | suit_bit<<12 | rank<<8 | prime[rank]."""
    c = c.to(torch.int64)
    r, s = c % 13, torch.div(c, 13, rounding_mode="floor")
    return ((1 << r) << 16) | (_TREYS_SUIT_BITS_T.to(c.device)[s] << 12) | (r << 8) | _PRIMES_T.to(c.device)[r]


def path_from_treys_t(t: torch.Tensor) -> torch.Tensor:
    """Treys card ints -> path encoding (int64), elementwise; malformed ints raise ValueError."""
    t = t.to(torch.int64)
    r, bit = (t >> 8) & 0xF, (t >> 12) & 0xF
    suit = torch.full_like(t, -1)
    for b, s in _SUIT_FROM_BIT.items():
        suit = torch.where(bit == b, torch.full_like(t, s), suit)
    rc = r.clamp(max=12)
    ok = (r <= 12) & (suit >= 0) & ((t >> 16) == (1 << rc)) & ((t & 0xFF) == _PRIMES_T.to(t.device)[rc])
    if not bool(ok.all()):
        raise ValueError("not a Treys card int")
    return suit * 13 + r


def situations_from_rank_major(hole: torch.Tensor, board: torch.Tensor) -> torch.Tensor:
    """Wire-format ints as GameBoard / Player publish them (rank-major, utils.py:82-92): hole [N,2], board [N,3..5] ->
    uint8 [N,8] situation rows.  CUDA tensors go through one fused kernel (holdem_pack_rank_major); values outside [0,52)
    become 0xFE there and the HS/HP kernels report the row as malformed.  CPU tensors are packed with torch (and validated
    here, raising ValueError)."""
    hole = hole.to(torch.int64)
    board = board.to(torch.int64)
    if hole.dim() != 2 or hole.shape[1] != 2 or board.dim() != 2 or board.shape[0] != hole.shape[0] \
            or not 3 <= board.shape[1] <= 5:
        raise ValueError("expected hole [N,2] and board [N,3..5]")
    if hole.is_cuda:
        from .api import pack_rank_major
        return pack_rank_major(hole, board)
    both = torch.cat([hole, board], dim=1)
    if both.numel() and (int(both.min()) < 0 or int(both.max()) > 51):
        raise ValueError("cards must be integers in 0..51")
    rows = torch.full((hole.shape[0], 8), 0xFF, dtype=torch.uint8)
    rows[:, :both.shape[1]] = path_from_rank_major_t(both).to(torch.uint8)
    return rows
