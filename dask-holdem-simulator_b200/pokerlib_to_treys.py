"""Drop-in replacement for the reference's pokerlib_to_treys.py (convert_to_treys, CalculateHandValue, device) plus
the two names the reference imports from the third-party `treys` package (Evaluator, Card), so that nothing on this
path needs `treys` installed: the evaluation runs in the batched CUDA evaluator (holdem_eval_batch) and returns
Treys' exact 1..7462 rank.

Reference call sites replaced: pokerlib_to_treys.py:12-13 (convert_to_treys), :15-21 (CalculateHandValue),
HandCalculations.py:231,238,268 (Evaluator().evaluate(board, hand)).
"""
import torch

from holdem_b200 import api as _api
from holdem_b200 import encoding as _enc

device = 'cuda:0' if torch.cuda.is_available() else 'cpu'


class Card:
    """The subset of treys.Card this path touches: the 32-bit card int and its accessors."""
    STR_RANKS = _enc.RANK_CHARS
    PRIMES = list(_enc.PRIMES)

    @staticmethod
    def new(string: str) -> int:
        return _enc.treys_from_str(string)

    @staticmethod
    def int_to_str(card_int: int) -> str:
        return _enc.str_from_path(_enc.path_from_treys(card_int))

    @staticmethod
    def get_rank_int(card_int: int) -> int:
        return (card_int >> 8) & 0xF

    @staticmethod
    def get_suit_int(card_int: int) -> int:
        return (card_int >> 12) & 0xF


class Evaluator:
    """treys.Evaluator look-alike: evaluate(hand, board) -> 1..7462, lower is better; 5, 6 or 7 cards in total.
    Constructing it is free (the tables live on the GPU, built once by holdem_init)."""

    MAX_RANKS = (10, 166, 322, 1599, 1609, 2467, 3325, 6185, 7462)
    CLASS_STRINGS = ("Straight Flush", "Four of a Kind", "Full House", "Flush", "Straight", "Three of a Kind",
                     "Two Pair", "Pair", "High Card")

    def evaluate(self, hand, board):
        cards = [_enc.path_from_treys(c) for c in list(hand) + list(board)]
        return evaluate_path_cards(cards)

    def get_rank_class(self, hand_rank: int) -> int:
        if not 1 <= hand_rank <= 7462:
            raise Exception("Inavlid hand rank, cannot return rank class")
        for cls, top in enumerate(self.MAX_RANKS, start=1):
            if hand_rank <= top:
                return cls
        raise AssertionError

    def class_to_string(self, class_int: int) -> str:
        return self.CLASS_STRINGS[class_int - 1]


def evaluate_path_cards(cards) -> int:
    """Treys rank of 5/6/7 cards in the path encoding (rank = c % 13, suit = c // 13) through the CUDA evaluator."""
    cards = [int(c) for c in cards]
    n = len(cards)
    if n not in (5, 6, 7):
        raise KeyError(n)  # what Evaluator.hand_size_map raises
    if not torch.cuda.is_available():
        raise RuntimeError("the Treys evaluator (B200 build) needs a CUDA device; there is no CPU fallback")
    row = torch.tensor([cards + [0xFF] * (8 - n)], dtype=torch.uint8, device='cuda')
    rank = int(_api.evaluate(row, n)[0].item())
    if rank == 0:
        raise ValueError(f"malformed hand (duplicate or unknown cards): {cards}")
    return rank


def convert_to_treys(cards):
    """poker.Card objects -> Treys card ints (TCard.new(card.rank.val + card.suit.value[1]) in the reference)."""
    return [_enc.treys_from_str(card.rank.val + card.suit.value[1]) for card in cards]


def CalculateHandValue(pokerlib_board, pokerlib_hand):
    """Normalised strength in [0,1): 1 - rank / 7462 (reference pokerlib_to_treys.py:15-21)."""
    converted_board = convert_to_treys(pokerlib_board)
    converted_hand = convert_to_treys(pokerlib_hand)
    hand_strength = Evaluator().evaluate(converted_board, converted_hand)
    return 1 - hand_strength / 7462
