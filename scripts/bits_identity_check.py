"""CPU check of the identity behind the flop kernel's bit-parallel undo half (kernels_hp_flop4.cu):
per rank pair pp, the total undo weight of a runout lane's step list
    [both cards in s | (a in s, rank of c outside s) | rank pair without a card of s],   step valid <=> its cards of s avoid the runout's
equals what the per-situation sets give
    BS & ~lo_t[q] & ~hi_t[q]  (weight 1)  +  XS[lo in s] & ~lo_t[q]  +  XS[hi in s] & ~hi_t[q]  (weight = unseen cards of the other
    rank outside s)  +  YS  (pairs of cards outside s),
for random unseen-card sets, suits, runout cards of the suit and step-list lengths (need = 2, 1, 0 cards of s in the holding).
usage: python scripts/bits_identity_check.py [trials]"""
import random
import sys


def tri(n):
    return n * (n + 1) // 2


PP = [(x, y) for y in range(13) for x in range(y + 1)]          # index tri(y) + x, x <= y
LO = [sum(1 << i for i, (x, y) in enumerate(PP) if x == r) for r in range(13)] + [0]
HI = [sum(1 << i for i, (x, y) in enumerate(PP) if y == r) for r in range(13)] + [0]


def check(trials=3000, seed=1):
    rng = random.Random(seed)
    for trial in range(trials):
        cards = rng.sample(range(52), 47)                        # unseen cards, card = 4 * rank + suit
        s = rng.randrange(4)
        in_s = sorted(c >> 2 for c in cards if (c & 3) == s)
        u = [sum(1 for c in cards if (c >> 2) == r) for r in range(13)]
        beta = [u[r] - (1 if r in in_s else 0) for r in range(13)]
        need = rng.choice([2, 1, 0])
        qbits = set(rng.sample(in_s, min(rng.choice([0, 1, 2]), len(in_s))))
        # the kernel's step lists
        w_steps = [0] * 91
        for i, a in enumerate(in_s):
            for b in in_s[i + 1:]:
                if a not in qbits and b not in qbits:
                    w_steps[tri(b) + a] += 1
        if need <= 1:
            for a in in_s:
                if a not in qbits:
                    for y in range(13):
                        w_steps[tri(max(a, y)) + min(a, y)] += beta[y]
        if need == 0:
            for i, (x, y) in enumerate(PP):
                w_steps[i] += beta[x] * beta[y] if x != y else beta[x] * (beta[x] - 1) // 2
        # the sets
        bs = sum(1 << i for i, (x, y) in enumerate(PP) if x != y and x in in_s and y in in_s)
        w1 = [beta[y] if x in in_s else 0 for (x, y) in PP]
        w2 = [beta[x] if (x != y and y in in_s) else 0 for (x, y) in PP]
        wy = [beta[x] * beta[y] if x != y else beta[x] * (beta[x] - 1) // 2 for (x, y) in PP]
        assert max(w1) <= 3 and max(w2) <= 3 and max(wy) <= 9    # two / four bit slices suffice
        r = sorted(qbits) + [13, 13]
        nlo, nhi = ~(LO[r[0]] | LO[r[1]]), ~(HI[r[0]] | HI[r[1]])
        w_sets = [0] * 91
        for i in range(91):
            bit = 1 << i
            if bs & nlo & nhi & bit:
                w_sets[i] += 1
            if need <= 1:
                w_sets[i] += (w1[i] if nlo & bit else 0) + (w2[i] if nhi & bit else 0)
            if need == 0:
                w_sets[i] += wy[i]
        assert w_steps == w_sets, (trial, need, sorted(qbits))
    return trials


if __name__ == "__main__":
    print("ok", check(int(sys.argv[1]) if len(sys.argv) > 1 else 3000), "trials")
