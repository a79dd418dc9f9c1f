// tables.h -- host-side construction of the evaluator lookup tables (built once per process, uploaded per device).
//
// What the tables replace: treys.LookupTable (flush_lookup / unsuited_lookup keyed by prime products) and the
// min-over-C(n,5) loops of treys.Evaluator._six/_seven, which the reference reaches through
// HandCalculations.py:229-240 and pokerlib_to_treys.py:15-21.  Nothing here follows Treys' construction: the
// 7462 five-card classes are produced by sorting explicit strength tuples (the standard poker total order, which
// is the order Treys numbers), the 6/7-card tables are a dynamic programme over rank multisets, and the sparse
// additive multiset keys are compressed with a hash-displace perfect hash sized for shared memory.
#pragma once
#include <cstdint>
#include <vector>

namespace holdem {

// Additive rank keys: the sum over the cards of a hand identifies the rank multiset uniquely for every hand of
// up to 7 cards with at most 4 cards per rank (checked by selftest()).  Largest 7-card sum: 7,825,759 (23 bits).
// Provenance: this particular key set is not ours -- it is the published non-flush "face key" constant set of the
// SpecialK / SKPokerEval family of public 7-card evaluators (K. J. Shackleton), used here as a constant only; nothing of the
// reference tree or of Treys is involved, and the one property relied on (unique sums) is re-verified at start-up.
constexpr uint32_t kRankKey[13] = {0,    1,     5,     22,     98,     453,    2031,
                                   8698, 22854, 83661, 262349, 636345, 1479181};

constexpr int kNumClasses = 7462;
constexpr int kFlushTableSize = 8192;
constexpr int kBoardMultisets = 455;     // C(15,3)
constexpr int kOwnMultisets = 6188;       // C(17,5)
constexpr int kBoardTPitch = 94;          // u16 per row of board_t
constexpr int kBoardTRow = 91 * kBoardTPitch + 6;   // u16 per board (17,120 bytes: a multiple of 16)
constexpr int kBoard4Multisets = 1820;    // C(16,4)
constexpr int kOwn6Multisets = 18564;     // C(18,6)
constexpr int kBoard4TRow = 91 * 16;      // u16 per four-card board (2,912 bytes)
constexpr int kMaxPairs = 1326;  // C(52,2), colex order: pair p of (i<j) has p = j(j-1)/2 + i

// Hash-displace perfect hash over the additive keys of one hand size:
//   bucket = (key * mul_b) >> (32 - bucket_bits)
//   slot   = (((key * mul_s) >> (32 - slot_bits)) + disp[bucket]) & (2^slot_bits - 1)
//   rank   = val[slot]
struct PerfectHash {
    uint32_t mul_b = 0, mul_s = 0;
    int bucket_bits = 0, slot_bits = 0;
    std::vector<uint16_t> disp;  // [2^bucket_bits]
    std::vector<uint16_t> val;   // [2^slot_bits], 0 = empty
    uint32_t lookup(uint32_t key) const {
        uint32_t b = (key * mul_b) >> (32 - bucket_bits);
        uint32_t s = (((key * mul_s) >> (32 - slot_bits)) + disp[b]) & ((1u << slot_bits) - 1);
        return val[s];
    }
};

struct Tables {
    // flush[mask]: rank of the best 5 cards of one suit whose ranks are the set bits of the 13-bit mask
    // (straight flush 1..10 or flush 323..1599) for popcount >= 5, else 0.
    std::vector<uint16_t> flush;
    // nf[n-5]: rank multiset of an n-card hand (n = 5, 6, 7) -> best five-card rank ignoring suits.
    PerfectHash nf[3];
    // pairs[p] = i | j << 8 for the p-th pair in colex order (prefix-stable: the first C(m,2) entries are the
    // pairs of {0..m-1} for every m).
    std::vector<uint16_t> pairs;
    // Static work list of the flop hand-potential kernel (kernels_hp_flop.cu): every (a,b,d), a<b<d-1, over the 47
    // unseen-card positions, grouped into warp tasks of 32 lanes that share the loop length d-b-1 (the lane then walks
    // c = b+1 .. d-1).  task_hdr[t] = k0 | iters << 8 (c starts at b+1+k0); flop_lanes[32 t + l] = a | b<<8 | d<<16 or
    // 0xFFFFFFFF for a padding lane.  Tasks are sorted by iters descending.
    std::vector<uint32_t> flop_task_hdr;
    std::vector<uint32_t> flop_lanes;
    // Same for the turn kernel (kernels_hp_turn.cu): every (a,b), a<b<45, over 46 positions, the lane walks c = b+1..45.
    // turn_lanes[32 t + l] = a | b<<8 or 0xFFFFFFFF.  nf3_ranks[e]: three ascending ranks, e = C(r0,1)+C(r1+1,2)+C(r2+2,3).
    std::vector<uint32_t> turn_task_hdr;
    std::vector<uint32_t> turn_lanes;
    std::vector<uint16_t> nf3_ranks;
    // Work lists of the reference-mode kernel (kernels_hp_ref.cu) for 46 and 45 unseen positions (turn / river boards;
    // the flop reuses flop_task_hdr / flop_lanes): same layout as the flop list.
    std::vector<uint32_t> ref_task_hdr[2];   // [0]: 46 positions, [1]: 45 positions
    std::vector<uint32_t> ref_lanes[2];
    // nf4_ranks[e] = the four ranks (4 bits each, ascending) of the e-th rank multiset, e = C(r0,1)+C(r1+1,2)+C(r2+2,3)+C(r3+3,4)
    std::vector<uint16_t> nf4_ranks;
    // Direct-indexed tables of the rank-multiset flop kernel, one contiguous row per situation (fetched with one bulk copy):
    //   board_t[b3][91][kBoardTPitch]  non-flush 7-card rank of board + rank pair q + rank pair p (symmetric in q, p),
    //                                  b3 = r0 + C(r1+1,2) + C(r2+2,3) over the sorted board ranks; 0 = impossible multiset
    //   board_opp5[b3][96]             non-flush 5-card rank of board + rank pair (the opponent on the flop)
    //   own_our7[m5][96]               non-flush 7-card rank of hole + board + rank pair, m5 = five-rank multiset index
    //   own_rank5[m5]                  non-flush 5-card rank of hole + board
    std::vector<uint16_t> board_t, board_opp5, own_our7, own_rank5;
    // The same for four-card boards (turn kernel, kernels_hp_turn4.cu), b4 = four-rank multiset index (nf4 numbering), m6 =
    // six-rank multiset index:
    //   board4_t[b4][91][16]   non-flush 7-card rank of board + river rank (column) + rank pair (row)
    //   board4_opp6[b4][96]    non-flush 6-card rank of board + rank pair
    //   own6_our7[m6][16]      non-flush 7-card rank of hole + board + river rank;  own6_rank6[m6]: of hole + board
    std::vector<uint16_t> board4_t, board4_opp6, own6_our7, own6_rank6;
    // Statistics kept for selftest().
    int n_keys[3] = {0, 0, 0};
};

const Tables &get_tables();  // thread-safe, builds on first use
int selftest(char *msg, int msg_len);  // 0 = ok

// Flat device image of the tables: one allocation, 16-byte aligned sections, offsets in bytes.
struct TableImage {
    std::vector<uint8_t> bytes;
    uint32_t off_flush = 0, off_pairs = 0, off_task_hdr = 0, off_lanes = 0, off_nf4 = 0;
    uint32_t off_board_t = 0, off_board_opp5 = 0, off_own_our7 = 0, off_own_rank5 = 0;
    uint32_t off_board4_t = 0, off_board4_opp6 = 0, off_own6_our7 = 0, off_own6_rank6 = 0;
    uint32_t n_tasks = 0;
    uint32_t off_turn_hdr = 0, off_turn_lanes = 0, off_nf3 = 0, n_turn_tasks = 0;
    uint32_t off_ref_hdr[2] = {0, 0}, off_ref_lanes[2] = {0, 0}, n_ref_tasks[2] = {0, 0};
    uint32_t off_disp[3] = {0, 0, 0}, off_val[3] = {0, 0, 0};
};
TableImage make_image(const Tables &t);

}  // namespace holdem
