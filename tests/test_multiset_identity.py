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


def test_reference_mode_identity_matches_literal_oracle(oracle):
    """The same identity for mode="reference" (scripts/proto_ref_multiset.py): literal 9-category ranker written as
    rank-count category + flush override, regular runouts plus the three kinds of runouts that repeat a hole card or one of
    the opponent's own cards; equals the literal oracle on monotone / two-tone / rainbow-paired boards."""
    import proto_ref_multiset as R
    assert R.check_ranker(trials=1500)
    sits = [([14, 1], [31, 32, 33]),      # testing/ExampleHS.py (golden counts of SURVEY 8c)
            ([0, 1], [2, 3, 4]),          # monotone board, we hold the straight flush
            ([12, 25], [38, 51, 0])]      # rainbow board, paired
    want = oracle.hp_batch(oracle.pack_situations(sits), "reference").astype(np.int64)
    assert want[0][3:12].tolist() == [334713, 148530, 172965, 70272, 265983, 130617, 27162, 16905, 104109]
    for k, (hole, board) in enumerate(sits):
        got, n_items = R.ref_flop_counts(hole, board)
        assert np.array_equal(got[:15], want[k][:15]), (hole, board, got, want[k])
        assert n_items < 1081 * 1176 // 4


def test_undo_half_step_lists_equal_rank_pair_sets():
    """The flop kernel's bit-parallel undo half: per rank pair, the weight of a runout lane's step list (validity included)
    equals what the per-situation sets BS / XS / YS and the static sets lo_t / hi_t give (scripts/bits_identity_check.py)."""
    import bits_identity_check as B
    assert B.check(1500, seed=7) == 1500
