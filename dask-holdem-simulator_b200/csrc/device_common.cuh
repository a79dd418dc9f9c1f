// device_common.cuh -- device-side evaluator primitives shared by every kernel of the path.
//
// Two rank semantics (SURVEY.md section 0 / include/holdem_b200.h):
//   * Treys rank 1..7462, lower is better (replaces treys.Evaluator.evaluate, reached from
//     HandCalculations.py:229-240).  A hand is summarised by (key, cards): key = sum of the additive rank keys
//     (tables.h), cards = 52-bit card set with bit (13*suit + rank) = the path's own card number.  Both are
//     additive over disjoint card groups, which is what lets the enumeration kernels build a 7-card hand as
//     board + opponent pair + runout with two integer adds.
//   * literal category 0..8, higher is better (replaces CalculateHandRankValue, HandCalculations.py:45-82),
//     computed arithmetically from nibble-packed rank counts so that multisets (duplicate cards) behave exactly
//     like the reference's Counter-based code.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace holdem {

struct NfHash {             // device view of tables.h::PerfectHash
    const uint16_t *disp;
    const uint16_t *val;
    uint32_t mul_b, mul_s;
    int bshift, sshift;     // 32 - bucket_bits, 32 - slot_bits
    uint32_t smask;
};

struct DevTables {          // pointers into the device-resident table image (global memory)
    const uint16_t *flush;  // [8192]
    const uint16_t *pairs;  // [1326] colex pairs, i | j << 8
    NfHash nf[3];           // 5, 6, 7 cards
    uint32_t disp_entries[3], val_entries[3];
    const uint32_t *flop_task_hdr;  // [n_flop_tasks] k0 | iters << 8
    const uint32_t *flop_lanes;     // [n_flop_tasks * 32] a | b<<8 | d<<16, 0xFFFFFFFF = padding lane
    const uint16_t *nf4_ranks;      // [1820] four ascending ranks, 4 bits each
    uint32_t n_flop_tasks;
    const uint32_t *turn_task_hdr;  // same layout for the turn kernel: lanes are a | b<<8
    const uint32_t *turn_lanes;
    const uint16_t *nf3_ranks;      // [455] three ascending ranks, 4 bits each
    uint32_t n_turn_tasks;
    const uint32_t *ref_task_hdr[2];  // reference-mode kernel, 46 / 45 unseen positions (47 uses the flop list)
    const uint32_t *ref_lanes[2];
    uint32_t n_ref_tasks[2];
    const uint16_t *board_t;          // [455][8560] non-flush rank of board + rank pair + rank pair (tables.h)
    const uint16_t *board_opp5;       // [455][96]
    const uint16_t *own_our7;         // [6188][96]
    const uint16_t *own_rank5;        // [6188]
    const uint16_t *board4_t;         // [1820][91][16] four-card boards (turn kernel)
    const uint16_t *board4_opp6;      // [1820][96]
    const uint16_t *own6_our7;        // [18564][16]
    const uint16_t *own6_rank6;       // [18564]
    const uint8_t *ref_t[3];          // HOLDEM_MODE_REFERENCE: [455 | 1820 | 6188][8384] category tables (kernels_hp_ref5.cu), device-built
};

// (provenance of the key set: tables.h)
static __constant__ uint32_t c_rank_key[13] = {0,    1,     5,     22,     98,     453,    2031,
                                        8698, 22854, 83661, 262349, 636345, 1479181};

// PTX shifts clamp the amount to the register width, so a card number outside [0,32) (or [32,64) for the high
// word) contributes no bit instead of invoking C++ undefined behaviour.
__device__ __forceinline__ uint32_t shl_clamp(uint32_t v, uint32_t s) {
    uint32_t r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(s));
    return r;
}

__device__ __forceinline__ uint64_t card_bit(uint32_t c) {
    return ((uint64_t)shl_clamp(1u, c - 32u) << 32) | shl_clamp(1u, c);
}

__device__ __forceinline__ uint32_t card_rank(uint32_t c) { return c - 13u * ((c * 79u) >> 10); }
__device__ __forceinline__ uint32_t card_suit(uint32_t c) { return (c * 79u) >> 10; }

// 13-bit rank mask of the suit holding at least `need` cards of the set, or 0.
__device__ __forceinline__ uint32_t flush_mask(uint64_t cards, int need) {
    uint32_t lo = (uint32_t)cards, hi = (uint32_t)(cards >> 32);
    uint32_t s0 = lo & 0x1FFFu;
    uint32_t s1 = (lo >> 13) & 0x1FFFu;
    uint32_t s2 = ((lo >> 26) | (hi << 6)) & 0x1FFFu;
    uint32_t s3 = (hi >> 7) & 0x1FFFu;
    uint32_t m = 0;
    m = __popc(s0) >= need ? s0 : m;
    m = __popc(s1) >= need ? s1 : m;
    m = __popc(s2) >= need ? s2 : m;
    m = __popc(s3) >= need ? s3 : m;
    return m;
}

__device__ __forceinline__ uint32_t nf_lookup(const NfHash &h, uint32_t key) {
    uint32_t b = (key * h.mul_b) >> h.bshift;
    uint32_t s = (((key * h.mul_s) >> h.sshift) + h.disp[b]) & h.smask;
    return h.val[s];
}

// Treys rank of a 5/6/7-card hand given its additive key and card set.  With five or more cards of one suit in
// at most seven cards no full house or quads can coexist, so the suited table alone decides.
__device__ __forceinline__ uint32_t treys_rank(const NfHash &h, const uint16_t *flush, uint32_t key,
                                               uint64_t cards) {
    uint32_t fm = flush_mask(cards, 5);
    return fm ? (uint32_t)flush[fm] : nf_lookup(h, key);
}

// ---- literal ranker (HandCalculations.py:24-82) ---------------------------------------------------------------
// cyclic five-run: is_straight()'s wrap-around list (:29-35) accepts A2345, KA234, QKA23, JQKA2 and TJQKA.
__device__ __forceinline__ bool cyclic_run5(uint32_t mask13) {
    uint32_t m = mask13 | (mask13 << 13);
    return (m & (m >> 1) & (m >> 2) & (m >> 3) & (m >> 4)) != 0;
}

struct LitHand {            // additive summary of a card multiset
    uint64_t rc;            // 13 nibbles: multiplicity of each rank            (value_counts, :57)
    uint64_t cards;         // OR of card bits
    uint64_t dup;           // cards present more than once
    uint32_t sc;            // 4 nibbles: multiplicity of each suit              (suit_counts, :56)
    uint32_t rmask;         // 13-bit set of ranks present
};

__device__ __forceinline__ void lit_add_card(LitHand &h, uint32_t c) {
    uint32_t s = card_suit(c), r = c - 13u * s;
    uint64_t bit = card_bit(c);
    h.rc += 1ull << (4 * r);
    h.sc += 1u << (4 * s);
    h.dup |= h.cards & bit;
    h.cards |= bit;
    h.rmask |= 1u << r;
}

// Combine disjoint-or-overlapping groups: counts add, card sets OR, duplicates are whatever two groups share.
__device__ __forceinline__ LitHand lit_merge(const LitHand &a, const LitHand &b) {
    LitHand h;
    h.rc = a.rc + b.rc;
    h.sc = a.sc + b.sc;
    h.dup = a.dup | b.dup | (a.cards & b.cards);
    h.cards = a.cards | b.cards;
    h.rmask = a.rmask | b.rmask;
    return h;
}

__device__ __forceinline__ uint32_t lit_category(const LitHand &h) {
    const uint64_t ONES = 0x1111111111111ull;
    // determine_flush (:39-43): a suit with >= 5 cards, multiplicity counted; at most one exists for n <= 9
    uint32_t fl = (h.sc + 0x3333u) & 0x8888u;
    if (fl) {
        uint32_t s = (uint32_t)(__ffs(fl) - 1) >> 2;
        uint32_t fm = (uint32_t)(h.cards >> (13 * s)) & 0x1FFFu;
        uint32_t fd = (uint32_t)(h.dup >> (13 * s)) & 0x1FFFu;
        if (fd == 0 && cyclic_run5(fm)) return 8;                       // :62-65
    }
    uint64_t b0 = h.rc & ONES, b1 = (h.rc >> 1) & ONES, b2 = (h.rc >> 2) & ONES, b3 = (h.rc >> 3) & ONES;
    uint64_t ge4 = b2 | b3;
    uint64_t ge5 = (b2 & (b1 | b0)) | b3;
    uint64_t eq3 = b0 & b1 & ~ge4;
    uint64_t eq2 = b1 & ~b0 & ~ge4;
    if (ge4 != 0 && ge5 == 0) return 7;                                  // :68  max == 4
    bool max3 = (ge4 == 0) && (eq3 != 0);
    if (max3 && eq2 != 0) return 6;                                      // :70
    if ((b1 | ge4) == 0 && cyclic_run5(h.rmask)) return 4;               // :72  all values distinct
    if (fl) return 5;                                                    // :74
    if (max3) return 3;                                                  // :76
    bool max2 = (ge4 == 0) && (eq3 == 0) && (eq2 != 0);
    if (max2 && __popcll(eq2) == 2) return 2;                            // :78
    if (max2) return 1;                                                  // :80
    return 0;                                                            // :82 (also max >= 5)
}

}  // namespace holdem
