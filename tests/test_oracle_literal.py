"""Pins the literal restatement (oracle/ehs_oracle.c) to golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py, reference HandCalculations.py imported behind stub modules)."""
import struct

import numpy as np
import pytest

from conftest import load_golden
from oracle import oracle as O
from oracle.reference_loader import reference_available


def f32hex(x):
    return struct.pack(">f", float(x)).hex()


def test_category_golden():
    g = load_golden("literal_category.json")["cases"]
    assert len(g) > 4000
    seen = set()
    for cards, value in g:
        assert O.lit_category(cards) == value, cards
        seen.add(value)
    assert seen == set(range(9))


def test_reference_unit_test_pins():
    # test_HandCalculations.py:17-18
    assert O.lit_category([0, 1, 10, 11, 12, 13, 14]) == 8
    pins = load_golden("helper_pins.json")
    # test_HandCalculations.py:28-45
    assert pins["remove_3"] == 49 and pins["remove_unknown"] == 52 and pins["comb_3"] == 3
    assert pins["opp_4known"] == 1128 and pins["runouts_flop"] == 1176
    for values, expect in pins["is_straight"]:
        assert O.lit_is_straight(values) == expect


def test_quirks():
    assert O.lit_category([0, 13, 1, 14, 2, 15, 30]) == 1      # three pairs -> "one pair"
    assert O.lit_category([0, 13, 26, 1, 14, 27, 41]) == 3     # two trips -> trips, not a full house
    assert O.lit_category([0, 14, 2, 16, 4, 13, 35]) == 1      # straight + pair -> not a straight
    assert O.lit_category([0, 1, 2, 3, 17, 5, 7]) == 4         # flush + unsuited straight -> 4, checked before 5
    assert O.lit_category([11, 12, 0, 14, 28, 45, 48]) == 4    # K A 2 3 4 wraps
    assert O.lit_category([0, 13, 26, 39, 0, 18, 33]) == 0   # five of one value (duplicate card)
    assert O.lit_category([0, 0, 0, 0, 0, 5, 7]) == 5          # duplicates count towards the flush


def test_hand_strength_golden():
    for c in load_golden("literal_hs.json")["cases"]:
        counts = O.hs_counts(c["our"], c["board"], "reference")
        assert O.hs_float(*counts).hex() == c["hs_hex"], c
    assert O.hs_float(*O.hs_counts([0, 13], [26, 27, 28, 29, 30], "reference")) == 0.5   # testing/ExampleHS.py:4-7


def test_hand_potential_golden():
    cases = load_golden("literal_hp.json")["cases"]
    assert {len(c["board"]) for c in cases} == {3, 4, 5}
    for c in cases:
        hp, tot = O.hp_counts(c["our"], c["board"], "reference")
        assert hp == c["HP"] and tot == c["HPTotal"], c
        ppot, npot = O.hp_floats(hp, tot)
        assert f32hex(ppot) == c["ppot_f32_hex"] and f32hex(npot) == c["npot_f32_hex"]
        assert list(O.hs_counts(c["our"], c["board"], "reference")) == tot
    # testing/ExampleHS.py:11-14 / artifacts/HandStrengthAndPotential.png
    ex = cases[0]
    assert ex["our"] == [14, 1] and ex["board"] == [31, 32, 33]
    assert round(ex["ppot"], 4) == 135.2782 and round(ex["npot"], 4) == 327.2654
    assert ex["HPTotal"] == [558, 397, 126] and sum(map(sum, ex["HP"])) == 1271256


@pytest.mark.skipif(not reference_available(), reason="reference tree not present (GPU box)")
def test_against_live_reference():
    """Re-validate against the imported reference when it is available (build container only)."""
    import random
    import torch
    from oracle.reference_loader import load_reference
    ref = load_reference()
    rng = random.Random(7)
    for _ in range(300):
        n = rng.choice([5, 6, 7, 8, 9])
        cards = [rng.randrange(52) for _ in range(n)]
        assert O.lit_category(cards) == ref.CalculateHandRankValue(cards[:2], cards[2:])
    for nb in (3, 4, 5):
        cards = rng.sample(range(52), 2 + nb)
        hs = ref.HandStrength(torch.tensor(cards[:2]), torch.tensor(cards[2:]))
        assert O.hs_float(*O.hs_counts(cards[:2], cards[2:], "reference")) == hs


def test_batch_driver_matches_scalar():
    sit = O.pack_situations([([14, 1], [31, 32, 33]), ([25, 50], [19, 30, 22, 37]), ([32, 8], [18, 50, 6, 39, 16])])
    out = O.hp_batch(sit, "reference")
    for row, (our, board) in zip(out, [([14, 1], [31, 32, 33]), ([25, 50], [19, 30, 22, 37]),
                                       ([32, 8], [18, 50, 6, 39, 16])]):
        hp, tot = O.hp_counts(our, board, "reference")
        assert row[3:12].tolist() == sum(hp, []) and row[12:15].tolist() == tot and row[15] == len(board)
        assert row[0:3].tolist() == tot
    hs = O.hs_batch(sit, "reference")
    assert np.array_equal(hs, out[:, 0:3])
