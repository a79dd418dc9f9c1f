"""Pure-Python restatement of the Treys evaluator -- the "Treys/Python path" CPU baseline of BASELINE.md section 3.

TEST / BASELINE INFRASTRUCTURE ONLY (never imported by the product).  `treys` itself is not installable here
(no network), so bench.py times this restatement -- same data structures as the package
(dict lookups keyed by prime products, min over itertools.combinations for 6/7 cards) -- and labels it as such.
Checked against oracle/treys_oracle.c in tests/test_oracle_treys.py.
"""
import itertools

PRIMES = [2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41]
STR_RANKS = "23456789TJQKA"
SUIT_BITS = {"s": 1, "h": 2, "d": 4, "c": 8}


def card_new(s: str) -> int:
    r = STR_RANKS.index(s[0])
    return ((1 << r) << 16) | (SUIT_BITS[s[1]] << 12) | (r << 8) | PRIMES[r]


def card_from_path(c: int) -> int:
    return card_new(STR_RANKS[c % 13] + "cdhs"[c // 13])


def _prime_product_from_rankbits(bits: int) -> int:
    p = 1
    for i in range(13):
        if bits & (1 << i):
            p *= PRIMES[i]
    return p


class PyEvaluator:
    def __init__(self):
        self.flush_lookup = {}
        self.unsuited_lookup = {}
        self._flushes()
        self._multiples()

    def _flushes(self):
        sfs = [7936, 3968, 1984, 992, 496, 248, 124, 62, 31, 4111]
        others = [m for m in range(8192) if bin(m).count("1") == 5 and m not in sfs]
        others.sort(reverse=True)
        for rank, m in enumerate(sfs, start=1):
            self.flush_lookup[_prime_product_from_rankbits(m)] = rank
            self.unsuited_lookup[_prime_product_from_rankbits(m)] = 1599 + rank
        for i, m in enumerate(others):
            self.flush_lookup[_prime_product_from_rankbits(m)] = 323 + i
            self.unsuited_lookup[_prime_product_from_rankbits(m)] = 6186 + i

    def _multiples(self):
        back = list(range(12, -1, -1))
        rank = 11
        for q in back:
            for k in back:
                if k != q:
                    self.unsuited_lookup[PRIMES[q] ** 4 * PRIMES[k]] = rank
                    rank += 1
        rank = 167
        for t in back:
            for p in back:
                if p != t:
                    self.unsuited_lookup[PRIMES[t] ** 3 * PRIMES[p] ** 2] = rank
                    rank += 1
        rank = 1610
        for t in back:
            for a, b in itertools.combinations([k for k in back if k != t], 2):
                self.unsuited_lookup[PRIMES[t] ** 3 * PRIMES[a] * PRIMES[b]] = rank
                rank += 1
        rank = 2468
        for p1, p2 in itertools.combinations(back, 2):
            for k in back:
                if k != p1 and k != p2:
                    self.unsuited_lookup[PRIMES[p1] ** 2 * PRIMES[p2] ** 2 * PRIMES[k]] = rank
                    rank += 1
        rank = 3326
        for p in back:
            for a, b, c in itertools.combinations([k for k in back if k != p], 3):
                self.unsuited_lookup[PRIMES[p] ** 2 * PRIMES[a] * PRIMES[b] * PRIMES[c]] = rank
                rank += 1

    def _five(self, cards):
        if cards[0] & cards[1] & cards[2] & cards[3] & cards[4] & 0xF000:
            bits = (cards[0] | cards[1] | cards[2] | cards[3] | cards[4]) >> 16
            return self.flush_lookup[_prime_product_from_rankbits(bits)]
        p = 1
        for c in cards:
            p *= c & 0xFF
        return self.unsuited_lookup[p]

    def evaluate(self, hand, board):
        cards = list(hand) + list(board)
        if len(cards) == 5:
            return self._five(cards)
        if len(cards) not in (6, 7):
            raise KeyError(len(cards))
        return min(self._five(c) for c in itertools.combinations(cards, 5))
