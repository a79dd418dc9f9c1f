"""CPU proof of the counting identity behind the production flop kernel (kernels_hp_flop4.cu): the rank-multiset sum with
card-level flush corrections, restated in Python (scripts/proto_flop_multiset.py), equals the oracle's plain enumeration
row for row -- on every board texture (monotone / two-tone / rainbow, hole cards in and out of the board's suits)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def test_rank_multiset_identity_matches_oracle(oracle):
    import proto_flop_multiset as P
    sits = [([14, 1], [31, 32, 33]),      # testing/ExampleHS.py: monotone board, hole cards outside the suit
            ([0, 1], [2, 3, 4]),          # monotone board, hole cards in the suit (we hold the flush)
            ([26, 27], [0, 2, 17]),       # two-tone board
            ([12, 25], [38, 51, 0]),      # rainbow board, paired
            ([13, 26], [0, 1, 15]),       # two-tone, we hold one card of each board suit
            ([39, 40], [41, 42, 0])]      # two-tone with four to a flush in hand + board
    want = oracle.hp_batch(oracle.pack_situations(sits), "treys").astype(np.int64)
    for k, (hole, board) in enumerate(sits):
        got, n_items = P.flop_counts(hole, board)
        assert np.array_equal(got, want[k]), (hole, board, got, want[k])
        assert n_items < 1_070_190 // 5       # the identity visits a small fraction of the (holding, runout) pairs


def test_rank_multiset_identity_turn_matches_oracle(oracle):
    """The one-river-card form of the identity (kernels_hp_turn4.cu) on boards with two, three and four cards of a suit."""
    import proto_flop_multiset as P
    sits = [([0, 1], [2, 3, 4, 5]),           # four to a flush on board, we hold the suit
            ([13, 26], [2, 3, 4, 18]),        # three of a suit on board
            ([0, 13], [2, 3, 15, 16]),        # two suits twice
            ([12, 25], [38, 51, 0, 13]),      # rainbow, two pairs on board
            ([39, 1], [41, 42, 43, 44])]      # four of a suit, one hole card in it
    want = oracle.hp_batch(oracle.pack_situations(sits), "treys").astype(np.int64)
    for k, (hole, board) in enumerate(sits):
        got, n_items = P.turn_counts(hole, board)
        assert np.array_equal(got, want[k]), (hole, board, got, want[k])
        assert n_items < 45540
