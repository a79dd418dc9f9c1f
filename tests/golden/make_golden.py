#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED
reference (/root/reference/HandCalculations.py) on CPU behind stub imports.

Run once in the build container (the reference tree is not available on the GPU
box):   python tests/golden/make_golden.py [--hp-flop 14 --hp-turn 5 --hp-river 5]

Outputs (committed):
  literal_category.json  CalculateHandRankValue (HandCalculations.py:45-82) on random
                         multisets n=5..9 (with and without duplicate cards) + quirk hands
  literal_hs.json        HandStrength (HandCalculations.py:157-176), boards of 3/4/5 cards
  literal_hp.json        HandPotential (HandCalculations.py:178-226): HP[3][3], HPTotal[3]
                         as integers (captured from the reference's own tensors) and the
                         returned float32 PPot/NPot as exact hex floats
  helper_pins.json       enumeration-helper pins (HandCalculations.py:131-154)
"""
import argparse
import json
import multiprocessing as mp
import os
import random
import struct
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
os.environ["CUDA_VISIBLE_DEVICES"] = ""


def f32_hex(x) -> str:
    return struct.pack(">f", float(x)).hex()


def deal(rng, n_board):
    cards = rng.sample(range(52), 2 + n_board)
    return cards[:2], cards[2:]


def _hp_worker(args):
    our, board = args
    import torch
    torch.set_num_threads(1)
    from oracle.reference_loader import load_reference, TorchZerosSpy
    ref = load_reference()
    spy = TorchZerosSpy(torch)
    ref.torch = spy  # only the module-global name; the file on disk is untouched
    t0 = time.time()
    ppot, npot = ref.HandPotential(torch.tensor(our), torch.tensor(board))
    dt = time.time() - t0
    hp, hptotal = spy.captured[0], spy.captured[1]
    assert tuple(hp.shape) == (3, 3) and tuple(hptotal.shape) == (3,)
    hs = ref.HandStrength(torch.tensor(our), torch.tensor(board))
    return {
        "our": our, "board": board,
        "HP": [[int(v) for v in row] for row in hp.tolist()],
        "HPTotal": [int(v) for v in hptotal.tolist()],
        "ppot_f32_hex": f32_hex(ppot.item()), "npot_f32_hex": f32_hex(npot.item()),
        "ppot": float(ppot.item()), "npot": float(npot.item()),
        "ppot_dtype": str(ppot.dtype), "ret_type": "list",
        "hs_hex": float(hs).hex(), "hs": float(hs),
        "seconds": round(dt, 2),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hp-flop", type=int, default=14)
    ap.add_argument("--hp-turn", type=int, default=5)
    ap.add_argument("--hp-river", type=int, default=5)
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    ap.add_argument("--skip-hp", action="store_true")
    a = ap.parse_args()

    import torch
    from oracle.reference_loader import load_reference
    ref = load_reference()

    # ---- CalculateHandRankValue ------------------------------------------------
    rng = random.Random(20251019)
    cat = []
    quirks = [  # SURVEY.md section 8 a3 / c: constructed hands that exercise every branch
        ([0, 1], [10, 11, 12, 13, 14]),            # reference's own test pin -> 8
        ([0, 13], [26, 39, 1, 14, 27]),            # quads + trips
        ([0, 13], [1, 14, 2, 15, 30]),             # three pairs
        ([0, 13], [26, 1, 14, 27, 40]),            # two trips
        ([0, 14], [2, 16, 4, 13, 30]),             # straight + pair in 7 cards
        ([0, 1], [2, 3, 4, 20, 30]),               # straight flush low
        ([0, 1], [2, 3, 17, 5, 7]),                # flush + non-flush straight
        ([11, 12], [0, 14, 28, 40, 45]),           # K A 2 3 4 wrap
        ([10, 11], [12, 13, 27, 40, 45]),          # Q K A 2 3 wrap
        ([0, 0], [0, 0, 0, 5, 7]),                 # five copies of one value
        ([0, 0], [13, 26, 39, 5, 7]),              # five of a value via duplicate
        ([0, 1], [2, 3, 4]), ([14, 1], [31, 32, 33]), ([0, 13], [26, 27, 28, 29, 30]),
        ([12, 25], [38, 51, 11]), ([0, 13], [26, 39, 12]),
        ([0, 1], [2, 3, 4, 5, 6, 7, 8]),           # nine cards
        ([5, 5], [5, 5, 18, 18, 31, 31, 44]),
    ]
    for our, board in quirks:
        cat.append({"our": our, "board": board, "value": ref.CalculateHandRankValue(our, board)})
    for _ in range(3000):
        n = rng.choice([5, 6, 7, 7, 7, 8, 9])
        if rng.random() < 0.5:
            cards = rng.sample(range(52), n)
        else:  # duplicates allowed (runouts colliding with hole cards in the reference)
            cards = [rng.randrange(52) for _ in range(n)]
            if rng.random() < 0.5:  # bias to few ranks/suits so rare branches are hit
                base = rng.sample(range(13), 3)
                cards = [rng.choice(base) + 13 * rng.randrange(4) for _ in range(n)]
        our, board = cards[:2], cards[2:]
        v = ref.CalculateHandRankValue(torch.tensor(our), torch.tensor(board))
        cat.append({"our": our, "board": board, "value": v})
    for _ in range(1500):  # suit-heavy hands for the flush / straight-flush branches
        n = rng.choice([5, 6, 7, 8, 9])
        s = rng.randrange(4)
        k = rng.randint(4, min(n, 9))
        cards = [13 * s + r for r in rng.sample(range(13), k)]
        while len(cards) < n:
            cards.append(rng.randrange(52))
        rng.shuffle(cards)
        our, board = cards[:2], cards[2:]
        cat.append({"our": our, "board": board, "value": ref.CalculateHandRankValue(our, board)})
    with open(os.path.join(HERE, "literal_category.json"), "w") as f:
        json.dump({"source": "reference HandCalculations.py:45-82 (unmodified, stub imports)",
                   "cases": [[c["our"] + c["board"], c["value"]] for c in cat]}, f)
    print("category cases:", len(cat), "histogram:",
          {v: sum(1 for c in cat if c["value"] == v) for v in range(9)})

    # ---- helper pins -------------------------------------------------------------
    pins = {
        "remove_3": len(ref.remove_cards_from_deck(torch.tensor([0, 1, 2]))),
        "remove_unknown": len(ref.remove_cards_from_deck(torch.tensor([52, 53]))),
        "comb_3": len(list(ref.generate_combinations(torch.tensor([0, 1, 2])))),
        "opp_4known": len(list(ref.generate_opponent_hands(torch.tensor([0, 1]), torch.tensor([2, 3])))),
        "runouts_flop": len(list(ref.generate_possible_turns_and_rivers(torch.tensor([0, 1, 2])))),
        "runouts_turn": len(list(ref.generate_possible_turns_and_rivers(torch.tensor([0, 1, 2, 3])))),
        "runouts_river": len(list(ref.generate_possible_turns_and_rivers(torch.tensor([0, 1, 2, 3, 4])))),
        "first_opp_hands": [t.tolist() for t in list(ref.generate_opponent_hands(
            torch.tensor([14, 1]), torch.tensor([31, 32, 33])))[:5]],
        "is_straight": [[v, ref.is_straight(v)] for v in
                        ([0, 1, 2, 3, 4], [12, 0, 1, 2, 3], [11, 12, 0, 1, 2], [8, 9, 10, 11, 12],
                         [0, 1, 2, 3, 3], [0, 2, 4, 6, 8], [0, 1, 2, 3, 4, 5, 6], [3, 4, 5, 6, 7, 11, 12])],
    }
    with open(os.path.join(HERE, "helper_pins.json"), "w") as f:
        json.dump(pins, f, indent=1)

    # ---- HandStrength --------------------------------------------------------------
    rng = random.Random(0)
    hs_cases = [([0, 13], [26, 27, 28, 29, 30]), ([0, 1], [2, 3, 4]), ([14, 1], [31, 32, 33])]
    for i in range(6):  # the SURVEY's seeded extras: Random(0).sample(range(52),7), boards 3,4,5,...
        cards = rng.sample(range(52), 7)
        nb = [3, 4, 5][i % 3]
        hs_cases.append((cards[:2], cards[2:2 + nb]))
    rng2 = random.Random(20251019)
    for i in range(120):
        hs_cases.append(deal(rng2, [3, 4, 5][i % 3]))
    hs_out = []
    for our, board in hs_cases:
        v = ref.HandStrength(torch.tensor(our), torch.tensor(board))
        hs_out.append({"our": our, "board": board, "hs_hex": float(v).hex(), "hs": v,
                       "type": type(v).__name__})
    with open(os.path.join(HERE, "literal_hs.json"), "w") as f:
        json.dump({"source": "reference HandCalculations.py:157-176 (unmodified)", "cases": hs_out}, f, indent=0)
    print("hs cases:", len(hs_out))

    # ---- HandPotential ---------------------------------------------------------------
    if a.skip_hp:
        return
    rng3 = random.Random(20251020)
    hp_cases = [([14, 1], [31, 32, 33])]                      # testing/ExampleHS.py:11-14
    hp_cases.append(([0, 1], [2, 3, 4]))                      # test_HandCalculations.py:56 (disabled test)
    while len(hp_cases) < a.hp_flop:
        hp_cases.append(deal(rng3, 3))
    for _ in range(a.hp_turn):
        hp_cases.append(deal(rng3, 4))
    for _ in range(a.hp_river):
        hp_cases.append(deal(rng3, 5))
    t0 = time.time()
    with mp.get_context("spawn").Pool(a.procs) as pool:
        hp_out = pool.map(_hp_worker, hp_cases, chunksize=1)
    print("hp cases:", len(hp_out), "wall %.1fs" % (time.time() - t0))
    with open(os.path.join(HERE, "literal_hp.json"), "w") as f:
        json.dump({"source": "reference HandCalculations.py:178-226 (unmodified; HP/HPTotal captured "
                             "from the reference's own torch.zeros tensors)",
                   "cases": hp_out}, f, indent=0)


if __name__ == "__main__":
    main()
