"""CPU prototype of the rank-multiset flop HS+HP algorithm (kernels_hp_flop4.cu), checked against the oracle.

Test infrastructure: validates the counting identity the CUDA kernel relies on before any GPU time is spent.
  HP[cls(P)][cmp(our7[Q], R(P,Q))] summed over disjoint (P,Q)
    = generic part  : sum over Q, over rank pairs pp of n_Q(pp) * [clsr(pp)] x cmp(our7[Q], NF(board+Q+pp))
    + flush part    : for every (P,Q) where board+Q has k>=3 cards of a suit s and P holds >= 5-k cards of s:
                      + [cls(P)] x cmp(our7[Q], flush(board+Q+P in s))  - [clsr(P)] x cmp(our7[Q], NF(board+Q+P))
"""
import itertools
import sys
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O


def evaluate_rows(rows, n):
    a = np.full((len(rows), 8), 255, dtype=np.uint8)
    a[:, :n] = np.array(rows, dtype=np.uint8)
    return O.treys_evaluate_batch(a, n).astype(np.int64)


def nonflush_cards(ranks):
    """cards with the given rank multiset and no five of a suit (same-rank cards get different suits)."""
    ranks = sorted(ranks)
    return [13 * (i % 4) + r for i, r in enumerate(ranks)]


def pp_index(x, y):
    return y * (y + 1) // 2 + x


def flop_counts(hole, board):
    known = set(hole) | set(board)
    U = sorted((c for c in range(52) if c not in known), key=lambda c: (c % 13, c // 13))
    n = len(U)
    assert n == 47
    rk = [c % 13 for c in U]
    su = [c // 13 for c in U]
    pairs = [(i, j) for i in range(n) for j in range(i + 1, n)]
    our7 = dict(zip(pairs, evaluate_rows([list(hole) + list(board) + [U[i], U[j]] for i, j in pairs], 7)))
    opp5 = dict(zip(pairs, evaluate_rows([list(board) + [U[i], U[j]] for i, j in pairs], 5)))
    ours_now = int(evaluate_rows([list(hole) + list(board)], 5)[0])
    cls = {p: (0 if ours_now < v else (1 if ours_now == v else 2)) for p, v in opp5.items()}
    tot = [sum(1 for p in pairs if cls[p] == i) for i in range(3)]
    brk = [c % 13 for c in board]
    u = [rk.count(r) for r in range(13)]
    # rank-level tables
    rank_pairs = [(x, y) for y in range(13) for x in range(y + 1)]
    assert all(pp_index(x, y) == k for k, (x, y) in enumerate(rank_pairs))
    # clsr: class from ranks only (non-flush 5-card rank of board + x + y)
    def valid(ms):
        return all(ms.count(r) <= 4 for r in set(ms))
    rows5, idx5 = [], []
    for k, (x, y) in enumerate(rank_pairs):
        ms = brk + [x, y]
        if valid(ms):
            rows5.append(nonflush_cards(ms)); idx5.append(k)
    v5 = evaluate_rows(rows5, 5)
    clsr = [None] * 91
    for k, v in zip(idx5, v5):
        clsr[k] = 0 if ours_now < v else (1 if ours_now == v else 2)
    # board monotone: the 5-card class from ranks must ignore the flush; nonflush_cards does that
    T = np.zeros((91, 91), dtype=np.int64)
    rows7, idx7 = [], []
    for qk, (q1, q2) in enumerate(rank_pairs):
        for pk, (x, y) in enumerate(rank_pairs):
            ms = brk + [q1, q2, x, y]
            if valid(ms):
                rows7.append(nonflush_cards(ms)); idx7.append((qk, pk))
    v7 = evaluate_rows(rows7, 7)
    for (qk, pk), v in zip(idx7, v7):
        T[qk, pk] = v
    lt = [0, 0, 0]
    gt = [0, 0, 0]
    # ---- generic part
    for (i, j) in pairs:
        q1, q2 = rk[i], rk[j]
        qk = pp_index(q1, q2)
        o = int(our7[(i, j)])
        chk = 0
        for pk, (x, y) in enumerate(rank_pairs):
            mx = u[x] - (x == q1) - (x == q2)
            my = u[y] - (y == q1) - (y == q2)
            nn = mx * my if x != y else mx * (mx - 1) // 2
            if nn <= 0:
                continue
            chk += nn
            R = T[qk, pk]
            assert R > 0
            c = clsr[pk]
            if o < R:
                lt[c] += nn
            elif o > R:
                gt[c] += nn
        assert chk == 990
    # ---- flush part
    flush_cache = {}
    def flush_rank(mask):
        if mask not in flush_cache:
            rs = [r for r in range(13) if mask >> r & 1]
            best = min(int(v) for v in evaluate_rows([list(c) for c in itertools.combinations(rs, 5)], 5))
            flush_cache[mask] = best
        return flush_cache[mask]
    n_items = 0
    Wc = {}
    for (i, j) in pairs:
        Wc[(i, j)] = Wc[(j, i)] = cls[(i, j)]
    for s in range(4):
        b_s = sum(1 for c in board if c // 13 == s)
        if b_s == 0:
            continue
        bmask = sum(1 << (c % 13) for c in board if c // 13 == s)
        Ls = [i for i in range(n) if su[i] == s]
        Lns = [i for i in range(n) if su[i] != s]
        both = [(a, b) for a, b in itertools.combinations(Ls, 2)]
        one = [tuple(sorted((a, z))) for a in Ls for z in Lns]
        none = [(a, b) for a, b in itertools.combinations(Lns, 2)]
        wns = [sum(1 for i in Lns if rk[i] == y) for y in range(13)]
        # per-suit aggregates used by the "one" and "none" blocks (kernels_hp_flop4.cu builds the same lists)
        histA = {a: [sum(1 for c in Lns if Wc[(a, c)] == k3) for k3 in range(3)] for a in Ls}
        wnn = [wns[x] * wns[y] if x != y else wns[x] * (wns[x] - 1) // 2 for (x, y) in rank_pairs]
        hist_none = [sum(w for w, c in zip(wnn, clsr) if c == k3 and w) for k3 in range(3)]
        hs_s = sum(1 for c in list(hole) + list(board) if c // 13 == s)
        histA_rank = {a: {y: [sum(1 for c in Lns if rk[c] == y and Wc[(a, c)] == k3) for k3 in range(3)]
                          for y in range(13)} for a in Ls}
        our7nf = {}
        for (i, j) in pairs:                      # our rank from ranks alone (no flush of ours): any suit assignment of
            cards = list(hole) + list(board) + [U[i], U[j]]                       # those ranks without five of a suit
            if max(sum(1 for c in cards if c // 13 == t) for t in range(4)) < 5:
                our7nf.setdefault(pp_index(rk[i], rk[j]), int(our7[(i, j)]))
        for qtype, qlist in ((2, both), (1, one), (0, none)):
            k = b_s + qtype
            if k < 3:
                continue
            need = 5 - k
            # lanes: (our, rank bits in s, rank pair index, weight, cards outside s, rank of the card outside s)
            rank_lanes = qtype != 2 and hs_s + qtype < 5      # none of these runouts gives us a flush: merge them by rank
            lanes = []
            if rank_lanes and qtype == 1:
                for a in Ls:
                    for rz in range(13):
                        if wns[rz]:
                            qk = pp_index(min(rk[a], rz), max(rk[a], rz))
                            lanes.append((our7nf[qk], 1 << rk[a], qk, wns[rz], [], rz))
            elif rank_lanes:
                for qk, (x, y) in enumerate(rank_pairs):
                    if wnn[qk]:
                        lanes.append((our7nf[qk], 0, qk, wnn[qk], [], None))
            else:
                for (i, j) in qlist:
                    qbits = sum(1 << rk[t] for t in (i, j) if su[t] == s)
                    lanes.append((int(our7[(i, j)]), qbits, pp_index(rk[i], rk[j]), 1,
                                  [t for t in (i, j) if su[t] != s], None))
            for (o, qbits, qk, wl, zs, rz) in lanes:
                fq = bmask | qbits
                dl = [0, 0, 0]
                dg = [0, 0, 0]
                # -- both cards of P in s: card level, disjointness = rank bits in s
                for (a, b) in both:
                    pbits = (1 << rk[a]) | (1 << rk[b])
                    if pbits & qbits:
                        continue
                    n_items += 1
                    Rf = flush_rank(fq | pbits)
                    pk = pp_index(rk[a], rk[b])
                    Rn = T[qk, pk]
                    ct, cr = cls[(a, b)], clsr[pk]
                    if o < Rf: dl[ct] += 1
                    elif o > Rf: dg[ct] += 1
                    if o < Rn: dl[cr] -= 1
                    elif o > Rn: dg[cr] -= 1
                if need <= 1:
                    assert len(zs) <= 1
                    for a in Ls:
                        if (1 << rk[a]) & qbits:
                            continue
                        n_items += 1
                        Rf = flush_rank(fq | (1 << rk[a]))
                        for c3 in range(3):
                            if o < Rf: dl[c3] += histA[a][c3]
                            elif o > Rf: dg[c3] += histA[a][c3]
                        for y in range(13):
                            if wns[y] <= 0:
                                continue
                            n_items += 1
                            pk = pp_index(min(rk[a], y), max(rk[a], y))
                            Rn = T[qk, pk]
                            cr = clsr[pk]
                            if o < Rn: dl[cr] -= wns[y]
                            elif o > Rn: dg[cr] -= wns[y]
                if need == 0:
                    assert not zs
                    Rf = flush_rank(fq)
                    for c3 in range(3):
                        if o < Rf: dl[c3] += hist_none[c3]
                        elif o > Rf: dg[c3] += hist_none[c3]
                    for pk in range(91):
                        if wnn[pk] <= 0:
                            continue
                        n_items += 1
                        Rn = T[qk, pk]
                        cr = clsr[pk]
                        if o < Rn: dl[cr] -= wnn[pk]
                        elif o > Rn: dg[cr] -= wnn[pk]
                for c3 in range(3):
                    lt[c3] += wl * dl[c3]
                    gt[c3] += wl * dg[c3]
                # fix-up: the runout's own card(s) outside s were counted among the c of {a, c}
                if need <= 1:
                    for a in Ls:
                        if (1 << rk[a]) & qbits:
                            continue
                        Rf = flush_rank(fq | (1 << rk[a]))
                        if rz is not None:            # merged lane: all wl cards z of rank rz at once
                            hz = histA_rank[a][rz]
                            pk = pp_index(min(rk[a], rz), max(rk[a], rz))
                            Rn = T[qk, pk]
                            cr = clsr[pk]
                            for c3 in range(3):
                                if o < Rf: lt[c3] -= hz[c3]
                                elif o > Rf: gt[c3] -= hz[c3]
                            if o < Rn: lt[cr] += wl
                            elif o > Rn: gt[cr] += wl
                        for z in zs:
                            ct = Wc[(a, z)]
                            pk = pp_index(min(rk[a], rk[z]), max(rk[a], rk[z]))
                            Rn = T[qk, pk]
                            cr = clsr[pk]
                            if o < Rf: lt[ct] -= 1
                            elif o > Rf: gt[ct] -= 1
                            if o < Rn: lt[cr] += 1
                            elif o > Rn: gt[cr] += 1
    row = np.zeros(16, dtype=np.int64)
    row[0:3] = tot
    for i in range(3):
        row[3 + 3 * i + 0] = lt[i]
        row[3 + 3 * i + 2] = gt[i]
        row[3 + 3 * i + 1] = 990 * tot[i] - lt[i] - gt[i]
    row[12:15] = tot
    row[15] = 3
    return row, n_items


def turn_counts(hole, board):
    """The same identity with four board cards and ONE river card q (kernels_hp_turn4.cu), in its plain form: generic sum over
    (river card, rank pair) plus one add / undo item per (river card, holding) that makes a flush with board + q."""
    known = set(hole) | set(board)
    U = sorted((c for c in range(52) if c not in known), key=lambda c: (c % 13, c // 13))
    n = len(U)
    assert n == 46
    rk = [c % 13 for c in U]
    su = [c // 13 for c in U]
    pairs = [(i, j) for i in range(n) for j in range(i + 1, n)]
    our7 = evaluate_rows([list(hole) + list(board) + [U[q]] for q in range(n)], 7)
    opp6 = dict(zip(pairs, evaluate_rows([list(board) + [U[i], U[j]] for i, j in pairs], 6)))
    ours_now = int(evaluate_rows([list(hole) + list(board)], 6)[0])
    cls = {p: (0 if ours_now < v else (1 if ours_now == v else 2)) for p, v in opp6.items()}
    tot = [sum(1 for p in pairs if cls[p] == i) for i in range(3)]
    brk = [c % 13 for c in board]
    u = [rk.count(r) for r in range(13)]
    rank_pairs = [(x, y) for y in range(13) for x in range(y + 1)]

    def valid(ms):
        return all(ms.count(r) <= 4 for r in set(ms))
    clsr = [None] * 91
    rows6 = [(k, nonflush_cards(brk + [x, y])) for k, (x, y) in enumerate(rank_pairs) if valid(brk + [x, y])]
    for (k, _), v in zip(rows6, evaluate_rows([r for _, r in rows6], 6)):
        clsr[k] = 0 if ours_now < v else (1 if ours_now == v else 2)
    T = np.zeros((13, 91), dtype=np.int64)
    rows7 = [((rq, pk), nonflush_cards(brk + [rq, x, y])) for rq in range(13) for pk, (x, y) in enumerate(rank_pairs)
             if valid(brk + [rq, x, y])]
    for ((rq, pk), _), v in zip(rows7, evaluate_rows([r for _, r in rows7], 7)):
        T[rq, pk] = v
    lt, gt = [0, 0, 0], [0, 0, 0]
    for q in range(n):                                  # generic part
        o = int(our7[q])
        for pk, (x, y) in enumerate(rank_pairs):
            mx, my = u[x] - (x == rk[q]), u[y] - (y == rk[q])
            nn = mx * my if x != y else mx * (mx - 1) // 2
            if nn <= 0:
                continue
            R, c = T[rk[q], pk], clsr[pk]
            if o < R: lt[c] += nn
            elif o > R: gt[c] += nn
    cache = {}

    def flush_rank(mask):
        if mask not in cache:
            rs = [r for r in range(13) if mask >> r & 1]
            cache[mask] = min(int(v) for v in evaluate_rows([list(c) for c in itertools.combinations(rs, 5)], 5))
        return cache[mask]
    n_items = 0
    for s in range(4):                                  # flush part
        b_s = sum(1 for c in board if c // 13 == s)
        bmask = sum(1 << (c % 13) for c in board if c // 13 == s)
        for q in range(n):
            k = b_s + (su[q] == s)
            if k < 3:
                continue
            fq = bmask | ((1 << rk[q]) if su[q] == s else 0)
            o = int(our7[q])
            for (a, b) in pairs:
                if q in (a, b) or (su[a] == s) + (su[b] == s) < 5 - k:
                    continue
                n_items += 1
                Rf = flush_rank(fq | sum(1 << rk[t] for t in (a, b) if su[t] == s))
                pk = pp_index(rk[a], rk[b])
                Rn, ct, cr = T[rk[q], pk], cls[(a, b)], clsr[pk]
                if o < Rf: lt[ct] += 1
                elif o > Rf: gt[ct] += 1
                if o < Rn: lt[cr] -= 1
                elif o > Rn: gt[cr] -= 1
    row = np.zeros(16, dtype=np.int64)
    row[0:3] = tot
    for i in range(3):
        row[3 + 3 * i + 0], row[3 + 3 * i + 2] = lt[i], gt[i]
        row[3 + 3 * i + 1] = 44 * tot[i] - lt[i] - gt[i]
    row[12:15] = tot
    row[15] = 4
    return row, n_items


def main():
    import random
    rng = random.Random(7)
    sits = []
    # crafted textures: monotone board (+ suited hole in / out of the suit), two-tone, rainbow, paired boards
    sits += [([14, 1], [31, 32, 33]), ([0, 1], [2, 3, 4]), ([0, 13], [5, 7, 9]), ([12, 11], [0, 2, 4]),
             ([26, 27], [0, 2, 17]), ([12, 25], [38, 51, 0]), ([5, 18], [31, 44, 6]), ([3, 4], [5, 6, 20]),
             ([12, 11], [10, 9, 8]), ([13, 26], [0, 1, 15])]
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 20):
        c = rng.sample(range(52), 5)
        sits.append((c[:2], c[2:]))
    packed = O.pack_situations(sits)
    want = O.hp_batch(packed, "treys").astype(np.int64)
    bad = 0
    for k, (h, b) in enumerate(sits):
        got, n_items = flop_counts(h, b)
        ok = np.array_equal(got, want[k])
        print(k, h, b, "items", n_items, "OK" if ok else f"MISMATCH\n got {got}\n want {want[k]}")
        bad += not ok
    print("mismatches:", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
