"""Drop-in replacement for the reference's HandCalculations.py (same public names, signatures and return types).

Put this directory in front of the reference on sys.path and `from HandCalculations import *` (GameBoard.py:3,
Player.py:4, testing/ExampleHS.py:1, test_HandCalculations.py:9) resolves here.  The counting loops of
HandStrength (reference HandCalculations.py:157-176), HandPotential (:178-226), CalculateHandRankValue (:45-82) and
the Treys evaluation behind evaluate_best_hand (:229-240) run in hand-written CUDA kernels for sm_100a through the
C ABI of include/holdem_b200.h.  There is no CPU fallback: importing works anywhere, calling needs a B200.

Rank semantics: environment variable HOLDEM_MODE, or set_mode():
  "reference" (default) the reference's code exactly as written (9-category ranker, 1176 runouts, quirks) -- this module is
              a drop-in, so by default every function returns what the original module returns, bit for bit, e.g.
              HandStrength([0,13],[26..30]) -> 0.5 and HandPotential([14,1],[31,32,33]) -> [135.2782, 327.2654].
  "treys"     opt-in: the canonical Effective-Hand-Strength enumeration on Treys ranks (what BASELINE.json grades and the
              batched API / bench.py use explicitly).  In this mode PPot/NPot are still the reference's unnormalised formula
              (:221-224); use HandPotentialCounts / EffectiveHandStrength or holdem_b200.normalised_potentials for
              probabilities.
The float formulas are the reference's own in both modes (:175 in Python floats, :221-224 in float32 tensors).
"""
import itertools
import os
import random  # noqa: F401  (re-exported by the reference's star import)
from collections import Counter  # noqa: F401

import torch

try:  # the reference re-exports its card/enum helpers; they need the third-party `poker` and `colorama`
    from utils import *  # type: ignore  # noqa: F401,F403
except Exception:  # pragma: no cover - optional glue, not on the hot path
    pass

from holdem_b200 import api as _api
from holdem_b200 import encoding as _enc
from pokerlib_to_treys import *  # noqa: F401,F403
from pokerlib_to_treys import convert_to_treys, Evaluator, Card  # noqa: F401

# Pre-flight check exactly like the reference (HandCalculations.py:22)
device = 'cuda' if torch.cuda.is_available() else 'cpu'

_MODE = os.environ.get("HOLDEM_MODE", "reference")
if _MODE not in ("treys", "reference"):
    raise ValueError(f"HOLDEM_MODE must be 'treys' or 'reference', got {_MODE!r}")


def set_mode(mode: str) -> None:
    """Select the rank semantics of HandStrength / HandPotential for this process."""
    global _MODE
    _api.mode_id(mode)
    _MODE = mode


def get_mode() -> str:
    return _MODE


def _as_list(cards):
    if isinstance(cards, torch.Tensor):
        cards = cards.tolist()
    return [int(c) for c in cards]


def _situation_tensor(ourcards, boardcards) -> torch.Tensor:
    our, board = _as_list(ourcards), _as_list(boardcards)
    if len(our) != 2 or not 3 <= len(board) <= 5:
        raise ValueError("expected 2 hole cards and a board of 3, 4 or 5 cards")
    if not torch.cuda.is_available():
        raise RuntimeError("HandCalculations (B200 build) needs a CUDA device; there is no CPU fallback")
    rows = _enc.pack_situations([our], [board])
    return torch.from_numpy(rows).to('cuda')


# ---- small host helpers kept for API compatibility (reference HandCalculations.py:24-43, :131-154) -----------------------
def is_straight(values):
    """True iff the values are pairwise distinct and five of them are consecutive on the 13-cycle
    (the reference's wrap-around: A2345, KA234, QKA23, JQKA2 and TJQKA all count)."""
    values = list(values)
    if len(set(values)) != len(values):
        return False
    ordered = sorted(values)
    doubled = ordered + [v + 13 for v in ordered]
    return any(all(doubled[i + k] == doubled[i] + k for k in range(1, 5))
               for i in range(len(doubled) - 4))


def determine_flush(suits, suit_counts):
    for suit, count in suit_counts.items():
        if count >= 5:
            return True, suit
    return False, None


def remove_cards_from_deck(cards_to_remove):
    """Full deck minus the given cards; unknown cards are reported and skipped like the reference (:135-136)."""
    deck = list(range(52))
    for card in cards_to_remove:
        card = int(card)
        if card in deck:
            deck.remove(card)
        else:
            print(f"Trying to remove card: {card} which is not in full_deck")
    return deck


def generate_combinations(cards):
    """All 2-card combinations as 2-element tensors on `device` (lazy)."""
    for a, b in itertools.combinations(list(cards), 2):
        yield torch.tensor([a, b], device=device)


def generate_opponent_hands(ourcards, boardcards):
    return generate_combinations(remove_cards_from_deck(_as_list(ourcards) + _as_list(boardcards)))


def generate_possible_turns_and_rivers(boardcards):
    return generate_combinations(remove_cards_from_deck(_as_list(boardcards)))


# ---- the hot path ----------------------------------------------------------------------------------------------------------
def CalculateHandRankValue(ourcards, boardcards):
    """Reference :45-82 -- category 0..8 (higher is better) of the card multiset ourcards + boardcards."""
    cards = _as_list(ourcards) + _as_list(boardcards)
    if not 2 <= len(cards) <= 9:
        raise ValueError("CalculateHandRankValue (B200 build) supports 2..9 cards")
    if not torch.cuda.is_available():
        raise RuntimeError("HandCalculations (B200 build) needs a CUDA device; there is no CPU fallback")
    row = torch.tensor([cards + [0xFF] * (16 - len(cards))], dtype=torch.uint8, device='cuda')
    value = int(_api.category(row, len(cards))[0].item())
    if value == 0xFF:
        raise ValueError(f"card outside 0..51 in {cards}")
    return value


def CalculateHandRank(ourcards, boardcards):
    """Reference :86-129.  Off the hot path and broken upstream: with five or more cards the reference raises
    TypeError inside its shadowed is_straight (:24-31 receives (Rank, Suit) tuples); with fewer it returns (0,).
    Kept importable with that behaviour."""
    if len(_as_list(ourcards) + _as_list(boardcards)) < 5:
        return (0,)
    raise TypeError("CalculateHandRank is broken in the reference for >= 5 cards (HandCalculations.py:102 calls the "
                    "list-based is_straight on (Rank, Suit) tuples); use CalculateHandRankValue or evaluate_best_hand")


def HandStrength(ourcards, boardcards):
    """Reference :157-176 -> Python float (ahead + tied/2) / (ahead + tied + behind)."""
    counts = _api.hand_strength_counts(_situation_tensor(ourcards, boardcards), _MODE)[0].tolist()
    if counts[0] < 0:
        raise ValueError("malformed situation (cards must be distinct and in 0..51)")
    ahead, tied, behind = counts
    handstrength = (ahead + tied / 2) / (ahead + tied + behind)
    return handstrength


def HandPotential(ourcards, boardcards):
    """Reference :178-226 -> [Ppot, Npot], two 0-dim float32 tensors on `device`, formula of :221-224 unchanged."""
    rows = _api.hand_potential_counts(_situation_tensor(ourcards, boardcards), _MODE)
    if int(rows[0, 15].item()) < 0:
        raise ValueError("malformed situation (cards must be distinct and in 0..51)")
    ppot, npot = _api.potentials_from_counts(rows)
    return [ppot[0], npot[0]]


def HandPotentialCounts(ourcards, boardcards):
    """Extension: the integer HP[3][3] matrix and HPTotal[3] the floats above are derived from."""
    row = _api.hand_potential_counts(_situation_tensor(ourcards, boardcards), _MODE)[0].tolist()
    if row[15] < 0:
        raise ValueError("malformed situation (cards must be distinct and in 0..51)")
    return [row[3:6], row[6:9], row[9:12]], row[12:15]


def EffectiveHandStrength(ourcards, boardcards):
    """Extension: HS*(1-NPot) + (1-HS)*PPot on the normalised potentials (parity unpinned, see api.ehs_from_counts)."""
    rows = _api.hand_potential_counts(_situation_tensor(ourcards, boardcards), _MODE)
    if int(rows[0, 15].item()) < 0:
        raise ValueError("malformed situation (cards must be distinct and in 0..51)")
    return float(_api.ehs_from_counts(rows)[0].item())


def evaluate_best_hand(player_cards, board_hand):
    """Reference :229-240 -- Treys rank (1..7462, lower is better) of poker.Card lists, returned as-is."""
    treys_player_cards = convert_to_treys(player_cards)
    treys_board_hand = convert_to_treys(board_hand)
    return evaluator.evaluate(treys_board_hand, treys_player_cards)


evaluator = Evaluator()
