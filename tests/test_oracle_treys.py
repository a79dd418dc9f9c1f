"""Pins the Treys restatement (oracle/treys_oracle.c) to external known answers.
The reference holds no test at the Treys boundary, so these facts are the pin."""
import multiprocessing as mp

import numpy as np

from oracle import oracle as O


def pc(s):
    return "cdhs".index(s[1]) * 13 + "23456789TJQKA".index(s[0])


def test_table_sizes_and_card_ints():
    L = O.lib()
    assert L.treys_flush_keys() == 1287
    assert L.treys_unsuited_keys() == 6175
    assert L.treys_card_new(11, b"d") == 0x08004B25   # Kd
    assert L.treys_card_new(3, b"s") == 0x00081307    # 5s
    assert L.treys_card_new(9, b"c") == 0x0200891D    # Jc
    assert O.treys_card_from_path(pc("Kd")) == 0x08004B25


def test_known_hands():
    ev = lambda *cards: O.treys_evaluate([pc(c) for c in cards])
    assert ev("Ah", "Kd", "Jc", "Qs", "Th") == 1600            # pokerlib_to_treys.py:24-28 (Treys README example)
    assert ev("As", "Ks", "Qs", "Js", "Ts") == 1                # royal flush
    assert ev("7s", "5d", "4c", "3h", "2s") == 7462             # pokerlib_to_treys.py:37-39
    assert ev("5s", "4s", "3s", "2s", "As") == 10               # steel wheel
    assert ev("Ac", "Ad", "Ah", "As", "Kc") == 11               # best quads
    assert ev("2c", "2d", "2h", "2s", "3c") == 166
    assert ev("Ac", "Ad", "Ah", "Ks", "Kc") == 167
    assert ev("As", "Ks", "Qs", "Js", "9s") == 323
    assert ev("Ac", "Ad", "Ah", "Ks", "Qc") == 1610
    assert ev("Ac", "Ad", "Kh", "Ks", "Qc") == 2468
    assert ev("Ac", "Ad", "Kh", "Qs", "Jc") == 3326
    assert ev("Ac", "Kd", "Qh", "Js", "9c") == 6186
    # 6 and 7 cards = best five
    assert ev("As", "Ks", "Qs", "Js", "Ts", "2c", "2d") == 1
    assert ev("Ah", "Kd", "Jc", "Qs", "Th", "2c") == 1600


def test_hand_size_dispatch():
    import pytest
    with pytest.raises(KeyError):
        O.treys_evaluate([0, 1, 2, 3])
    with pytest.raises(KeyError):
        O.treys_evaluate(list(range(8)))


def test_five_card_classes():
    import itertools
    rows = np.array([list(c) + [255, 255, 255] for c in itertools.combinations(range(52), 5)], dtype=np.uint8)
    assert rows.shape[0] == 2598960
    r = O.treys_evaluate_batch(rows, 5)
    assert len(np.unique(r)) == 7462 and r.min() == 1 and r.max() == 7462


def _slice(c0):
    return O.treys_exhaustive7(c0, c0 + 1)


def test_exhaustive_seven_card_histogram():
    """All C(52,7) = 133,784,560 hands: published 7-card class frequencies + order-independent checksums."""
    with mp.get_context("fork").Pool() as pool:
        parts = pool.map(_slice, range(46), chunksize=1)
    hist = sum(p[0] for p in parts)
    s = sum(p[1] for p in parts)
    s2 = sum(p[2] for p in parts)
    assert int(hist.sum()) == 133784560
    bounds = [0, 10, 166, 322, 1599, 1609, 2467, 3325, 6185, 7462]
    per_class = [int(hist[bounds[i] + 1:bounds[i + 1] + 1].sum()) for i in range(9)]
    assert per_class == [41584, 224848, 3473184, 4047644, 6180020, 6461620, 31433400, 58627800, 23294460]
    assert int(hist[1]) == 4324 and int(hist[10]) == 4140 and int(hist[11]) == 4052 and int(hist[1600]) == 747980
    assert int(hist[7462]) == 0 and int((hist > 0).sum()) == 4824
    assert s == 547965983972 and s2 == 2665153084455708
    # the whole 7463-bin histogram pinned by one hash: the GPU sweep (tests/test_gpu_eval.py) and any later table change must
    # reproduce it -- the product's independent table builder and this restatement agree bin by bin
    import json
    import os
    import numpy as np
    import holdem_b200 as hb
    want = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "gpu_hashes.json")))["exhaustive7_rank_histogram"]
    assert f"{hb.rows_hash_host(hist.astype(np.uint64).view(np.uint32).reshape(-1, 2)):016x}" == want
