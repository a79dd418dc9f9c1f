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
