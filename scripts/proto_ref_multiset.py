"""CPU prototype: the rank-multiset counting identity for mode="reference" (the reference's own loops, HandCalculations.py:178-226
with the 9-category ranker of :45-82), checked against the literal oracle.

Test infrastructure / design proof for the next reference-mode kernel: mode="treys" already runs on this identity
(kernels_hp_flop4.cu, scripts/proto_flop_multiset.py); the reference's variant differs in three ways, all handled here:

  * the ranker is the literal category 0..8 (higher = better), where "straight" is tested before "flush" and needs all n
    values distinct, a repeated card counts twice towards pairs and flushes and forbids the straight flush, and five of a
    value (four real cards + a repeat) falls through to 0;
  * runouts are ALL pairs of the deck minus the board (1176 on the flop), so besides the 990 regular runouts per holding P there are
      (i)   Q = {h, x}: one of our hole cards + an unseen card outside P              2 x 45
      (ii)  Q = {e, x}: one of the opponent's own cards + an unseen card outside P     2 x 45
      (iii) Q inside P u hole                                                          6
  * the five-card class of a holding uses the same literal ranker.

For every type the count is
    sum over Q, over rank groups g of the free opponent cards:  n_Q(g) * [clsr(g)] x cmp(our(Q), NF(rank counts))
  + for every (P, Q) whose opponent multiset holds >= 5 cards of a suit (repeats counted):
        + [cls(P)] x cmp(our(Q), cat(board + Q + P))  -  [clsr(P)] x cmp(our(Q), NF(rank counts))
with NF = the category from the rank counts alone and the flush override  cat = 8 if straight flush, else NF if NF in
{7, 6, 4}, else 5.  Type (iii) is summed directly (6 items per holding).
"""
import itertools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O


def cyc(mask):
    m = mask | (mask << 13)
    return (m & (m >> 1) & (m >> 2) & (m >> 3) & (m >> 4)) != 0


def cat_nf(ranks):
    """category from the rank counts alone (HandCalculations.py:56-82 without the flush tests)."""
    cnt = [0] * 13
    mask = 0
    for r in ranks:
        cnt[r] += 1
        mask |= 1 << r
    maxc = max(cnt)
    if maxc == 4:
        return 7
    if maxc == 3 and 2 in cnt:
        return 6
    if maxc == 1 and cyc(mask):
        return 4
    if maxc == 3:
        return 3
    if maxc == 2 and cnt.count(2) == 2:
        return 2
    if maxc == 2:
        return 1
    return 0


def flush_suit(cards):
    sc = [0] * 4
    for c in cards:
        sc[c // 13] += 1
    for s in range(4):
        if sc[s] >= 5:
            return s
    return None


def cat_full(cards):
    """the literal category of a card multiset, written as NF + flush override."""
    nf = cat_nf([c % 13 for c in cards])
    s = flush_suit(cards)
    if s is None:
        return nf
    vals = [c % 13 for c in cards if c // 13 == s]
    if len(set(vals)) == len(vals) and cyc(sum(1 << v for v in vals)):
        return 8
    return nf if nf in (7, 6, 4) else 5


def cmp_idx(our, opp):
    return 0 if our > opp else (1 if our == opp else 2)


def ref_flop_counts(hole, board):
    hole, board = list(hole), list(board)
    known = set(hole) | set(board)
    U = [c for c in range(52) if c not in known]
    n = len(U)
    assert n == 47
    rk = [c % 13 for c in U]
    su = [c // 13 for c in U]
    brk = [c % 13 for c in board]
    u = [rk.count(r) for r in range(13)]
    pairs = [(i, j) for i in range(n) for j in range(i + 1, n)]
    ours_now = cat_full(hole + board)
    cls = {}
    for (i, j) in pairs:
        cls[(i, j)] = cls[(j, i)] = cmp_idx(ours_now, cat_full(board + [U[i], U[j]]))
    tot = [sum(1 for p in pairs if cls[p] == k) for k in range(3)]

    def clsr(x, y):
        return cmp_idx(ours_now, cat_nf(brk + [x, y]))

    HP = np.zeros((3, 3), dtype=np.int64)
    n_items = 0

    def flush_possible(q_cards, s):
        """cards of suit s the opponent gets from board + runout (repeats counted)."""
        return sum(1 for c in board if c // 13 == s) + sum(1 for c in q_cards if c // 13 == s)

    # ---- regular runouts and type (i): the opponent's seven cards are distinct, P ranges over pairs of `free` ---------------
    def pair_block(q_cards, our_cards, free):
        nonlocal n_items
        o = cat_full(our_cards)
        qr = [c % 13 for c in q_cards]
        uf = [0] * 13
        for i in free:
            uf[rk[i]] += 1
        for y in range(13):
            for x in range(y + 1):
                nn = uf[x] * uf[y] if x != y else uf[x] * (uf[x] - 1) // 2
                if nn:
                    HP[clsr(x, y)][cmp_idx(o, cat_nf(brk + qr + [x, y]))] += nn
        for s in range(4):
            k = flush_possible(q_cards, s)
            if k < 3:
                continue
            Ls = [i for i in free if su[i] == s]
            others = [i for i in free if su[i] != s]
            items = list(itertools.combinations(Ls, 2))
            if k >= 4:
                items += [(a, z) for a in Ls for z in others]
            if k >= 5:
                items += list(itertools.combinations(others, 2))
            for (a, b) in items:
                n_items += 1
                opp = board + q_cards + [U[a], U[b]]
                HP[cls[(a, b)]][cmp_idx(o, cat_full(opp))] += 1
                HP[clsr(rk[a], rk[b])][cmp_idx(o, cat_nf([c % 13 for c in opp]))] -= 1

    for (i, j) in pairs:                                   # regular: Q inside the unseen cards, P disjoint from it
        pair_block([U[i], U[j]], hole + board + [U[i], U[j]], [t for t in range(n) if t != i and t != j])
    for h in hole:                                         # (i): Q = {our hole card, x}
        for x in range(n):
            pair_block([h, U[x]], hole + board + [h, U[x]], [t for t in range(n) if t != x])

    # ---- type (ii): Q = {e, x} with e the opponent's own card, P = {e, v} ----------------------------------------------------
    for e in range(n):
        for x in range(n):
            if x == e:
                continue
            q_cards = [U[e], U[x]]
            o = cat_full(hole + board + q_cards)
            free = [t for t in range(n) if t != e and t != x]
            uf = [0] * 13
            for t in free:
                uf[rk[t]] += 1
            base = brk + [rk[e], rk[x], rk[e]]
            for y in range(13):
                if uf[y]:
                    HP[clsr(min(rk[e], y), max(rk[e], y))][cmp_idx(o, cat_nf(base + [y]))] += uf[y]
            for s in range(4):
                # board + Q + the repeated e: cards of s before v is added
                k = flush_possible(q_cards, s) + (1 if su[e] == s else 0)
                if k < 4:
                    continue
                vs = [t for t in free if su[t] == s] if k == 4 else free
                for v in vs:
                    n_items += 1
                    opp = board + q_cards + [U[e], U[v]]
                    HP[cls[(e, v)]][cmp_idx(o, cat_full(opp))] += 1
                    HP[clsr(min(rk[e], rk[v]), max(rk[e], rk[v]))][cmp_idx(o, cat_nf([c % 13 for c in opp]))] -= 1

    # ---- type (iii): both runout cards are special, six per holding -----------------------------------------------------------
    our_hh = cat_full(hole + board + hole)
    for (i, j) in pairs:
        P = [U[i], U[j]]
        c3 = cls[(i, j)]
        HP[c3][cmp_idx(cat_full(hole + board + P), cat_full(board + P + P))] += 1
        HP[c3][cmp_idx(our_hh, cat_full(board + P + hole))] += 1
        for p in P:
            for h in hole:
                HP[c3][cmp_idx(cat_full(hole + board + [p, h]), cat_full(board + P + [p, h]))] += 1

    row = np.zeros(16, dtype=np.int64)
    row[0:3] = tot
    row[3:12] = HP.reshape(-1)
    row[12:15] = tot
    row[15] = 3
    return row, n_items


def check_ranker(trials=20000, seed=3):
    """cat_full (NF + flush override) equals the oracle's literal ranker on random multisets of 5..9 cards."""
    rng = np.random.default_rng(seed)
    for ncards in range(5, 10):
        rows = rng.integers(0, 52, size=(trials, 9), dtype=np.uint8)
        rows[::3] = (rng.integers(0, 2, size=(len(rows[::3]), 9)) * 13 + rng.integers(0, 7, size=(len(rows[::3]), 9))).astype(np.uint8)
        want = O.lit_category_batch(rows, ncards)
        for r, w in zip(rows, want):
            assert cat_full([int(c) for c in r[:ncards]]) == int(w), (r[:ncards], w)
    return True


def main():
    import random
    check_ranker(3000)
    rng = random.Random(11)
    sits = [([14, 1], [31, 32, 33]), ([0, 1], [2, 3, 4]), ([26, 27], [0, 2, 17]), ([12, 25], [38, 51, 0])]
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
        c = rng.sample(range(52), 5)
        sits.append((c[:2], c[2:]))
    want = O.hp_batch(O.pack_situations(sits), "reference").astype(np.int64)
    bad = 0
    for k, (h, b) in enumerate(sits):
        got, n_items = ref_flop_counts(h, b)
        ok = np.array_equal(got[:15], want[k][:15])
        print(k, h, b, "flush items", n_items, "of", 1081 * 1176, "OK" if ok else f"MISMATCH\n got {got}\n want {want[k]}")
        bad += not ok
    print("mismatches:", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
