"""Canonical (treys-mode) count oracle: regression vectors + structural properties.  Not reference-pinned (the
reference never runs this variant): anchored on the Treys restatement's known-answer tests."""
from oracle import oracle as O


def test_hs_regression_vectors():
    assert O.hs_counts([0, 13], [26, 27, 28, 29, 30], "treys") == (0, 946, 44)
    assert O.hs_counts([0, 1], [2, 3, 4], "treys") == (1080, 0, 1)
    assert O.hs_counts([14, 1], [31, 32, 33], "treys") == (564, 1, 516)
    assert O.hs_counts([24, 48], [26, 2, 16], "treys") == (476, 9, 596)
    assert O.hs_counts([25, 50], [19, 30, 22, 37], "treys") == (971, 6, 58)
    assert O.hs_counts([32, 8], [18, 50, 6, 39, 16], "treys") == (717, 6, 267)


def test_hp_flop_regression_and_cache_equivalence():
    hp, tot = O.hp_counts([14, 1], [31, 32, 33], "treys", cached=False)
    assert hp == [[276753, 35790, 245817], [0, 658, 332], [28713, 33025, 449102]]
    assert tot == [564, 1, 516]
    assert sum(map(sum, hp)) == 1081 * 990
    assert O.hp_counts([14, 1], [31, 32, 33], "treys", cached=True) == (hp, tot)
    ppot, npot = O.hp_floats(hp, tot)
    assert float(ppot) == 87.47679138183594 and float(npot) == 467.04071044921875


def test_hp_flop_4set_cache_equals_plain_loop():
    """The 4-set cache (178,365 opponent evaluations per situation) against the plain loop (1,070,190) on every
    board texture, including our own flush draws; and the batch driver's two cache settings against each other."""
    import numpy as np
    cases = [([14, 1], [31, 32, 33]),            # monotone board, we hold none of it
             ([0, 13], [26, 27, 28]),
             ([5, 7], [1, 3, 9]),                # monotone, we hold two of the suit
             ([5, 20], [1, 3, 30]),              # two-tone
             ([12, 25], [0, 14, 28]),            # rainbow, paired with nothing
             ([12, 25], [11, 24, 37]),           # trips on board
             ([24, 48], [26, 2, 16])]
    for our, board in cases:
        plain = O.hp_counts(our, board, "treys", cached=False)
        assert O.hp_counts(our, board, "treys", cached="4set") == plain, (our, board)
    sit = O.pack_situations(cases)
    assert np.array_equal(O.hp_batch(sit, "treys", cache4=True), O.hp_batch(sit, "treys", cache4=False))
    assert np.array_equal(O.hp_batch_parallel(sit, "treys", procs=2), O.hp_batch(sit, "treys"))


def test_hp_turn_and_river():
    our, board = [25, 50], [19, 30, 22, 37]
    hp, tot = O.hp_counts(our, board, "treys", cached=False)
    assert O.hp_counts(our, board, "treys", cached=True) == (hp, tot)
    assert sum(tot) == 1035 and sum(map(sum, hp)) == 1035 * 44
    assert tuple(tot) == O.hs_counts(our, board, "treys")
    for i in range(3):
        assert sum(hp[i]) == 44 * tot[i]
    hp5, tot5 = O.hp_counts([32, 8], [18, 50, 6, 39, 16], "treys")
    assert sum(map(sum, hp5)) == 0 and sum(tot5) == 990


def test_float_formulas_edge_cases():
    import math
    assert O.hs_float(1080, 1, 0) == 0.9995374653098983
    ppot, npot = O.hp_floats([[0] * 3] * 3, [0, 0, 990])
    assert float(ppot) == 0.0 and math.isnan(float(npot))
    ppot, npot = O.hp_floats([[5, 0, 0], [0, 0, 0], [0, 0, 0]], [0, 0, 0])
    assert math.isnan(float(ppot))
    ppot, npot = O.hp_floats([[0, 0, 7], [0, 0, 0], [0, 0, 0]], [0, 0, 3])
    assert math.isinf(float(npot))
