/*
 * oracle/treys_oracle.c -- CPU restatement of the Treys hand evaluator.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (dask-holdem-simulator_b200/)
 * may include, link or call this file; it exists to CHECK the CUDA path (tests/,
 * __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs).
 *
 * Provenance.  The reference reaches the evaluator only through the third-party
 * package `treys` (requirements.txt:3, unpinned; README.md:62-63 links
 * github.com/ihendley/treys; PyPI line 0.1.x), whose source is NOT under
 * /root/reference and is not installable here (no network).  Call sites in the
 * reference:  HandCalculations.py:5 (import), :231,:238 (Evaluator().evaluate(board, hand)),
 * :268 (module-level Evaluator()), pokerlib_to_treys.py:12-13 (Card.new(rank+suit)),
 * :19-21 (1 - rank/7462).  The reference holds no test that pins a value at this
 * boundary => PARITY UNPINNED BY THE REFERENCE at the Treys
 * boundary; this restatement is pinned instead by the package's published
 * algorithm and by external known-answer facts (tests/test_oracle_treys.py):
 * 7462 distinct classes, 1287 flush keys / 6175 unsuited keys, README example
 * Ah Kd Jc + Qs Th -> 1600, royal flush -> 1, 7-5-4-3-2 offsuit -> 7462
 * (pokerlib_to_treys.py:24-28,37-39), card ints Kd=0x08004B25 5s=0x00081307
 * Jc=0x0200891D, and the exhaustive 7-card class histogram / rank checksum.
 *
 * Published algorithm restated (treys card.py / lookup.py / evaluator.py; deuces lineage):
 *   card int   = (1<<rank)<<16 | suit<<12 | rank<<8 | prime[rank]
 *                rank 0..12 for "23456789TJQKA"; suit s=1 h=2 d=4 c=8
 *   LookupTable.flushes()      : straight flushes 1..10, other flushes 323..1599 keyed by the
 *                                 prime product of the rank bits; same bit patterns unsuited give
 *                                 straights 1600..1609 and high cards 6186..7462
 *   LookupTable.multiples()    : quads 11.., full houses 167.., trips 1610.., two pair 2468..,
 *                                 pairs 3326.., keyed by prime product, iterated from ace down
 *   Evaluator._five            : flush test on the suit nibble, then one table lookup
 *   Evaluator._six / _seven    : minimum of _five over all C(n,5) subsets
 *   Evaluator.evaluate(a, b)   : dispatch on len(a+b) in {5,6,7}
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static const int PRIMES[13] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41};

enum {
    MAX_STRAIGHT_FLUSH = 10, MAX_FOUR_OF_A_KIND = 166, MAX_FULL_HOUSE = 322, MAX_FLUSH = 1599,
    MAX_STRAIGHT = 1609, MAX_THREE_OF_A_KIND = 2467, MAX_TWO_PAIR = 3325, MAX_PAIR = 6185,
    MAX_HIGH_CARD = 7462
};

/* ---- tiny open-addressing map: prime product -> rank --------------------------------- */
#define MAP_BITS 15
#define MAP_SIZE (1u << MAP_BITS)
typedef struct { uint32_t key[MAP_SIZE]; uint16_t val[MAP_SIZE]; int count; } map_t;

static map_t flush_lookup, unsuited_lookup;
static int tables_ready = 0;

static uint32_t map_slot(uint32_t key) { return (key * 2654435761u) >> (32 - MAP_BITS); }

static void map_put(map_t *m, uint32_t key, int val) {
    uint32_t s = map_slot(key);
    while (m->key[s] != 0 && m->key[s] != key) s = (s + 1) & (MAP_SIZE - 1);
    if (m->key[s] == 0) m->count++;
    m->key[s] = key;
    m->val[s] = (uint16_t)val;
}

static int map_get(const map_t *m, uint32_t key) {
    uint32_t s = map_slot(key);
    while (m->key[s] != 0) {
        if (m->key[s] == key) return m->val[s];
        s = (s + 1) & (MAP_SIZE - 1);
    }
    return -1; /* KeyError in the Python original */
}

/* Card.prime_product_from_rankbits */
static uint32_t prime_product_from_rankbits(int rankbits) {
    uint32_t product = 1;
    for (int i = 0; i < 13; i++)
        if (rankbits & (1 << i)) product *= (uint32_t)PRIMES[i];
    return product;
}

/* LookupTable.get_lexographically_next_bit_sequence (bit hack) */
static int next_bit_sequence(int bits) {
    int t = (bits | (bits - 1)) + 1;
    return t | ((((t & -t) / (bits & -bits)) >> 1) - 1);
}

static void build_flushes(void) {
    static const int straight_flushes[10] = {7936, 3968, 1984, 992, 496, 248, 124, 62, 31, 4111};
    int flushes[1287];
    int nfl = 0;
    int bits = 31; /* 0b11111 */
    /* 1277 = number of high cards; iterate 1277 + 10 - 1 further sequences after the first */
    for (int i = 0; i < 1277 + 10 - 1; i++) {
        bits = next_bit_sequence(bits);
        int is_sf = 0;
        for (int j = 0; j < 10; j++)
            if (bits == straight_flushes[j]) is_sf = 1; /* original tests `not f ^ sf` */
        if (!is_sf) flushes[nfl++] = bits;
    }
    /* flushes.reverse(): strongest (largest bit pattern) first */
    int rank = 1;
    for (int j = 0; j < 10; j++)
        map_put(&flush_lookup, prime_product_from_rankbits(straight_flushes[j]), rank++);
    rank = MAX_FULL_HOUSE + 1;
    for (int j = nfl - 1; j >= 0; j--)
        map_put(&flush_lookup, prime_product_from_rankbits(flushes[j]), rank++);
    /* straight_and_highcards: same patterns, unsuited */
    rank = MAX_FLUSH + 1;
    for (int j = 0; j < 10; j++)
        map_put(&unsuited_lookup, prime_product_from_rankbits(straight_flushes[j]), rank++);
    rank = MAX_PAIR + 1;
    for (int j = nfl - 1; j >= 0; j--)
        map_put(&unsuited_lookup, prime_product_from_rankbits(flushes[j]), rank++);
}

static void build_multiples(void) {
    int rank;
    /* four of a kind */
    rank = MAX_STRAIGHT_FLUSH + 1;
    for (int i = 12; i >= 0; i--)
        for (int k = 12; k >= 0; k--) {
            if (k == i) continue;
            uint32_t p = (uint32_t)PRIMES[i];
            map_put(&unsuited_lookup, p * p * p * p * (uint32_t)PRIMES[k], rank++);
        }
    /* full house */
    rank = MAX_FOUR_OF_A_KIND + 1;
    for (int i = 12; i >= 0; i--)
        for (int pr = 12; pr >= 0; pr--) {
            if (pr == i) continue;
            uint32_t p = (uint32_t)PRIMES[i], q = (uint32_t)PRIMES[pr];
            map_put(&unsuited_lookup, p * p * p * q * q, rank++);
        }
    /* three of a kind: itertools.combinations(kickers descending, 2) */
    rank = MAX_STRAIGHT + 1;
    for (int r = 12; r >= 0; r--)
        for (int c1 = 12; c1 >= 0; c1--) {
            if (c1 == r) continue;
            for (int c2 = c1 - 1; c2 >= 0; c2--) {
                if (c2 == r) continue;
                uint32_t p = (uint32_t)PRIMES[r];
                map_put(&unsuited_lookup, p * p * p * (uint32_t)PRIMES[c1] * (uint32_t)PRIMES[c2], rank++);
            }
        }
    /* two pair: combinations(ranks descending, 2) x kicker descending */
    rank = MAX_THREE_OF_A_KIND + 1;
    for (int p1 = 12; p1 >= 0; p1--)
        for (int p2 = p1 - 1; p2 >= 0; p2--)
            for (int k = 12; k >= 0; k--) {
                if (k == p1 || k == p2) continue;
                uint32_t a = (uint32_t)PRIMES[p1], b = (uint32_t)PRIMES[p2];
                map_put(&unsuited_lookup, a * a * b * b * (uint32_t)PRIMES[k], rank++);
            }
    /* pair: pair rank descending x combinations(kickers descending, 3) */
    rank = MAX_TWO_PAIR + 1;
    for (int pr = 12; pr >= 0; pr--)
        for (int k1 = 12; k1 >= 0; k1--) {
            if (k1 == pr) continue;
            for (int k2 = k1 - 1; k2 >= 0; k2--) {
                if (k2 == pr) continue;
                for (int k3 = k2 - 1; k3 >= 0; k3--) {
                    if (k3 == pr) continue;
                    uint32_t p = (uint32_t)PRIMES[pr];
                    map_put(&unsuited_lookup,
                            p * p * (uint32_t)PRIMES[k1] * (uint32_t)PRIMES[k2] * (uint32_t)PRIMES[k3], rank++);
                }
            }
        }
}

void treys_init(void) {
    if (tables_ready) return;
    memset(&flush_lookup, 0, sizeof flush_lookup);
    memset(&unsuited_lookup, 0, sizeof unsuited_lookup);
    build_flushes();
    build_multiples();
    tables_ready = 1;
}

int treys_flush_keys(void) { treys_init(); return flush_lookup.count; }
int treys_unsuited_keys(void) { treys_init(); return unsuited_lookup.count; }

/* Card.new: rank_int 0..12, suit_char in "shdc" */
uint32_t treys_card_new(int rank_int, char suit_char) {
    int suit_int = suit_char == 's' ? 1 : suit_char == 'h' ? 2 : suit_char == 'd' ? 4 : 8;
    return ((1u << rank_int) << 16) | ((uint32_t)suit_int << 12) | ((uint32_t)rank_int << 8) |
           (uint32_t)PRIMES[rank_int];
}

/* Path encoding (HandCalculations.py:53-54, testing/ExampleHS.py:4-5,11-12):
 * card in [0,52), rank = card % 13 (0 = deuce), suit = card / 13 (0 c, 1 d, 2 h, 3 s). */
uint32_t treys_card_from_path(int card) { return treys_card_new(card % 13, "cdhs"[card / 13]); }

/* Evaluator._five */
int treys_five(const uint32_t *c) {
    if (c[0] & c[1] & c[2] & c[3] & c[4] & 0xF000u) {
        int hand_or = (int)((c[0] | c[1] | c[2] | c[3] | c[4]) >> 16);
        return map_get(&flush_lookup, prime_product_from_rankbits(hand_or));
    }
    uint32_t prime = 1;
    for (int i = 0; i < 5; i++) prime *= (c[i] & 0xFFu); /* Card.prime_product_from_hand */
    return map_get(&unsuited_lookup, prime);
}

/* Evaluator.evaluate on treys card ints: dispatch on n, 6/7 = min over all 5-subsets */
int treys_evaluate_ints(const uint32_t *cards, int n) {
    treys_init();
    if (n == 5) return treys_five(cards);
    if (n != 6 && n != 7) return -1; /* hand_size_map KeyError */
    int best = MAX_HIGH_CARD;
    uint32_t h[5];
    for (int a = 0; a < n; a++)
        for (int b = a + 1; b < n; b++)
            for (int c = b + 1; c < n; c++)
                for (int d = c + 1; d < n; d++)
                    for (int e = d + 1; e < n; e++) {
                        h[0] = cards[a]; h[1] = cards[b]; h[2] = cards[c]; h[3] = cards[d]; h[4] = cards[e];
                        int s = treys_five(h);
                        if (s < best) best = s;
                    }
    return best;
}

/* evaluate on path-encoded cards (suit-major 0..51) */
int treys_evaluate(const int *cards, int n) {
    uint32_t t[7];
    if (n < 5 || n > 7) return -1;
    for (int i = 0; i < n; i++) t[i] = treys_card_from_path(cards[i]);
    return treys_evaluate_ints(t, n);
}

/* batch helper: packed uint8 [N,8] rows (first n columns used) -> uint16 ranks */
void treys_evaluate_batch(const uint8_t *rows, int n, int64_t count, uint16_t *out) {
    treys_init();
    for (int64_t i = 0; i < count; i++) {
        int c[7];
        for (int j = 0; j < n; j++) c[j] = rows[i * 8 + j];
        out[i] = (uint16_t)treys_evaluate(c, n);
    }
}

/* Exhaustive sweep over all C(52,7) hands (lexicographic c0<...<c6), restricted to first-card
 * range [c0_lo, c0_hi) so callers can thread it.  hist[7463], sum, sumsq are accumulated. */
void treys_exhaustive7(int c0_lo, int c0_hi, uint64_t *hist, uint64_t *sum, uint64_t *sumsq) {
    treys_init();
    uint32_t t[52];
    for (int i = 0; i < 52; i++) t[i] = treys_card_from_path(i);
    uint32_t h[7];
    for (int a = c0_lo; a < c0_hi && a < 46; a++)
        for (int b = a + 1; b < 47; b++)
            for (int c = b + 1; c < 48; c++)
                for (int d = c + 1; d < 49; d++)
                    for (int e = d + 1; e < 50; e++)
                        for (int f = e + 1; f < 51; f++)
                            for (int g = f + 1; g < 52; g++) {
                                h[0] = t[a]; h[1] = t[b]; h[2] = t[c]; h[3] = t[d];
                                h[4] = t[e]; h[5] = t[f]; h[6] = t[g];
                                int r = treys_evaluate_ints(h, 7);
                                hist[r]++;
                                *sum += (uint64_t)r;
                                *sumsq += (uint64_t)r * (uint64_t)r;
                            }
}
