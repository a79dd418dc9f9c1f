// kernels_hp_flop4.cu -- flop hand strength + hand potential in Treys mode, rank-multiset formulation ("v4").
// One persistent CTA of 960 threads per SM = six independent groups of 160 threads, each working on its own situation
// (named barriers per group): the groups share the static tables and hide each other's latencies.  sm_100a only.
//
// What is counted (HandCalculations.py:178-226 with rank = Treys evaluate, SURVEY.md section 8a "canonical" loop):
//     HP[cls(P)][cmp(our7[Q], R(P,Q))] += 1     for every opponent holding P and runout Q, P and Q disjoint pairs of the
//                                                47 unseen cards (1081 x 990 = 1,070,190 pairs per situation)
// with cls(P) = ahead/tied/behind on the flop, our7[Q] = our rank on board+Q, R(P,Q) = the opponent's rank on board+Q+P.
//
// How (bit-identical counts, ~25x fewer compare steps than visiting every pair; scripts/proto_flop_multiset.py is the CPU
// proof of the identity, tests/test_gpu_hshp.py cross-checks this kernel against the plain enumeration and the oracle):
//   R(P,Q) is the flush-table rank when board+Q+P holds five cards of a suit and otherwise depends on the RANKS of P and Q
//   only.  So
//     sum over (P,Q)  =  GENERIC part: for every Q and every rank pair pp = (x<=y):
//                            n_Q(pp) * [clsr(pp)] x cmp(our7[Q], NF(board + ranks(Q) + pp))
//                        with n_Q(pp) = number of unseen pairs of those ranks disjoint from Q (a product of per-rank
//                        counts) and clsr = the flop class from ranks alone; all Q of one rank pair that give us no flush
//                        share our7, so this is 91 merged runouts (+ one entry per runout that makes OUR flush) x 91 rank
//                        pairs per situation;
//                     +  FLUSH part: for every (P,Q) such that board+Q has k >= 3 cards of a suit s and P holds >= 5-k
//                        cards of s (about 8 K pairs on a rainbow flop, 40 K two-tone, 140-210 K monotone; 7 K / 17 K /
//                        65 K items after the mergers below):
//                            + [cls(P)]  x cmp(our7[Q], flush[ranks of board+Q+P in s])
//                            - [clsr(P)] x cmp(our7[Q], NF(board + ranks(Q) + ranks(P)))        (undo the generic term)
//                        Holdings with BOTH cards in s are visited one by one.  Holdings {a in s, c not in s} share the
//                        flush rank of a, so per a they cost one compare against the class histogram of {a,c} over all c
//                        plus one rank-level undo step per rank of c; holdings without a card of s (only when board+Q is
//                        already a flush) cost one compare plus one undo step per rank pair.  Runouts get the same
//                        treatment from their side: a runout card outside s matters through its rank only, so those
//                        runouts are merged into weighted lanes (also when they complete OUR flush, which is true for all of
//                        them or none and does not depend on that card); a runout card outside s that coincides with c is
//                        taken out again by a short per-lane fix-up loop.
//   Our own flushes need no special case: our7[Q] is evaluated per card pair Q.
//
// Shared memory (~150 KB per CTA): static flush[8192] u16 (index bits folded, m ^ m >> 6, so that the bank depends on ten
// rank bits) and the colex pair list, then per group (~33 KB): T[91][91] u16 (symmetric) = non-flush rank of board + rank
// pair q + rank pair p, our7nf[91] and opp5nf[91] -- three rows of direct-indexed tables (tables.h: board_t, own_our7,
// board_opp5; indexed by the board's / the five known cards' rank multiset) that one thread fetches with cp.async.bulk
// onto mbarriers at the start of the situation, so the kernel performs no hash lookups at all; the runouts that give us a
// flush; PD[<=320] 16-byte step descriptors of the flush part, per suit ordered
// [both cards in s | one card a in s | undo steps (a, rank of c) | undo steps (rank pair), no card in s].
// Flush part mapping: a warp task = 32 runouts Q of one (suit, cards-of-Q-in-suit) group; lanes own Q (its rank, suit mask and
// column of T live in registers).
//   * add half: in closed form per lane when no runout of the warp gives us a flush or better (every add part is then a loss for
//     us: inclusion-exclusion over per-card class histograms); otherwise the warp walks the steps that have an add part with one
//     broadcast 16-byte descriptor load each (suited-table gather, two compares, two predicated multiply-adds of packed per-class
//     weights);
//   * undo half (kBits, round 2): the lane's runout shares its our7 and its column of T with merged runout e of the generic
//     part, so the compare outcomes it needs are the ones the generic tasks have just computed.  The generic tasks leave them
//     behind as two 91-bit sets per entry (LB / GB: our7 < T, our7 > T), the group meets at a barrier, and the undo half of a
//     lane's WHOLE step list becomes a weighted population count: the step weights by rank pair are per-situation sets built with
//     warp ballots (BS: both cards in s, weight 1; XS: {a in s, c outside s}, bit slices of the number of unseen cards of c's rank
//     outside s, by which of the two ranks is a's; YS: no card in s, bit slices of the pair count), the flop classes are three more
//     sets (CL), and a step is valid when its cards of s avoid the runout's, i.e. outside the static sets lo_t / hi_t of the
//     runout's ranks in s.  ~75 instructions per lane for a 66-step list, ~330 for a 210-step one, instead of 10 per step.
//     The runouts that complete OUR flush have generic entries of their own (suited our7): one per pair of cards inside the suit,
//     and -- because a runout card outside the suit changes neither our flush nor anything but a rank -- one per (card of the
//     suit, rank of the other card) / per rank pair with a weight; at most 224 entries, all with compare bits, so every task group
//     takes this path and the step-by-step undo half only exists in the kBits = false instantiation (holdem_hp_flop_variant 7).
//   19.6 K -> 13 K warp instructions per situation, 43.5 -> 60 M situations/s together with: the set-up phase spread over the five
//   warps, rank-merged lanes for the runouts that complete our flush, unequal parts of the generic chunks, the next claim made by
//   the first warp that runs out of flush tasks.
#include "device_common.cuh"
#include "launch.h"
#include "multiset_common.cuh"

namespace holdem {

namespace {

// Group shape (threads per situation group, groups per CTA).  Measured on B200, 65,536 random flops: 6 x 160 -> 39.0 M
// situations/s, 7 x 128 (one more situation in flight per SM) -> 38.8 M: the kernel is bound by instruction issue and
// shared-memory wavefronts inside the task phase, not by the latency of the per-situation set-up phases.
#ifndef HOLDEM_F4_CLAIM
#define HOLDEM_F4_CLAIM 1
#endif
#ifndef HOLDEM_F4_UNROLL
#define HOLDEM_F4_UNROLL 2
#endif
#ifndef HOLDEM_F4_GT
#define HOLDEM_F4_GT 160
#define HOLDEM_F4_G 6
#endif
constexpr int kF4Unroll = HOLDEM_F4_UNROLL;      // step unroll of the flush part's add loop
constexpr int kGroupThreads = HOLDEM_F4_GT;
constexpr int kGroups = HOLDEM_F4_G;
constexpr int kThreads4 = kGroupThreads * kGroups;
constexpr int kPairs4 = 1081;
constexpr int kPP = 91;            // rank pairs x <= y, index y(y+1)/2 + x
constexpr int kTPitch = 94;        // u16 units (tables.h kBoardTPitch)
constexpr int kBoardTRow4 = kPP * kTPitch + 6;   // u16 per board (tables.h kBoardTRow): 17,120 bytes
constexpr int kSeg = 64;           // longest step segment of a flush task (packed 10-bit counters hold 1023)
constexpr int kPD = 320;           // step descriptors per situation: at most 3 x 66 (rainbow) or 45 + 14 x 11 + 91 + 66
constexpr int kMaxGroups = 12;
constexpr int kProfSlots = 8;
constexpr int kExt = 224;           // entries of the runouts that complete our flush: C(10,2) | C(9,2) + 13 x 9 | C(8,2) + 13 x 8 + 91
constexpr int kBitEntries = 320;   // entries of the generic part (their compare bits are kept): 91 + kExt at most
constexpr int kClaim = HOLDEM_F4_CLAIM;           // situations a group claims from the global work counter at a time

struct Flop4Group {                    // one situation in flight
    uint4 PD[kPD];                     // x = 2 * (rank bits of the step's cards in s) << 16 | 2 * fold(those bits),
                                       // y = weight << 10 cls(P) (add part), z = weight << 10 clsr(pp) (undo part),
                                       // w = byte offset of row pp of T; 10-bit fields per class
    uint16_t T[kBoardTRow4];           // [pp][qp] (symmetric): non-flush rank of board + rank pair + rank pair, bulk-copied row
                                       // of DevTables::board_t
    uint16_t our7nf[96];               // non-flush rank of hole + board + rank pair (row of DevTables::own_our7)
    uint16_t opp5nf[96];               // non-flush rank of board + rank pair (row of DevTables::board_opp5)
    unsigned long long bar[2];         // mbarriers: [0] the two small rows, [1] T
    uint32_t ext[kExt];                // runouts that give us a flush: rank pair | our7 << 8 | number of merged runouts << 21
    uint32_t clsw[96];                 // per rank pair: 1 << 10 clsr, 0 when the pair cannot exist with this board
    uint32_t qflush[96];               // per rank pair: runouts moved to ext[]
    uint32_t cnt[2][12];               // lt[3], gt[3], class totals[3]; [situation parity]: the rows are written out of one
                                       // while the next situation's set-up clears the other
    uint32_t bmask[4];                 // board rank mask per suit
    uint32_t g_first[kMaxGroups + 1];  // first task id of each flush group; [n_groups] = total
    uint32_t g_nq[kMaxGroups], g_np[kMaxGroups], g_nseg[kMaxGroups], g_seglen[kMaxGroups], g_pdbase[kMaxGroups];
    uint32_t g_sq[kMaxGroups];         // suit | qtype << 2
    uint32_t pd_base[4], pd_len[4];    // PD section per suit
    uint32_t smask[4];                 // rank mask of the unseen cards per suit
    uint32_t hist_none[4];             // holdings without a card of the board's suit, by flop class (monotone boards)
    uint32_t Wc0[4][16];               // per suit and card a of it: class histogram (10-bit fields) of the holdings {a, b}, b in s
    uint32_t Hs[4][16];                // ... of the holdings {a, c}, c outside s (the add weight of the "card a in s" step)
    uint32_t Wt0[4], Ht[4];            // the sums of Wc0 / 2 and of Hs over the suit's cards
    uint32_t HA[16 * 13];              // monotone boards: class histogram (10-bit fields) of {a, c} over the c of one rank outside s
    uint32_t upack[2];                 // unseen cards per rank, 3 bits each (64-bit)
    uint32_t n_groups, pd_total, next_task, ours_now, n_ext;
    uint32_t next_claim[2];            // first work index of the claim being worked on / of the next one (alternating)
    uint32_t claim_rows[2][kClaim];    // the rows behind those work indices (resolved through the work order when the claim is made)
    uint32_t prof[kProfSlots];
    uint8_t code[48];                  // unseen cards, rank*4+suit, ascending
    uint8_t Ls[4][16];                 // positions of the unseen cards of suit s
    uint8_t Lns[4][48];                // positions of the unseen cards not of suit s
    uint8_t ns[4];
    uint8_t u[16];                     // unseen cards per rank
    uint8_t clsr[96];                  // per rank pair: flop class from ranks alone (3 = impossible pair)
    // ---- bit-parallel undo half (kBits): 91-bit sets of rank pairs, bit pp of word pp / 32 ----------------------------------
    uint32_t LB[3][kBitEntries], GB[3][kBitEntries];   // per entry e of the generic part (91 merged runouts, then the runouts that
                                       // complete our flush when they are the pairs of ten cards): { pp : our7[e] < T[e][pp] } and
                                       // { ... > ... }, written by the generic tasks (the two halves share word 1: atomicOr onto zeroed words)
    uint32_t CL[3][3];                 // [class][word]: the rank pairs of each flop class
    uint32_t BS[4][3];                 // [suit][word]: holdings with both cards in s, by rank pair
    uint32_t XS[2][2][3];              // [which card is in s: lo / hi][weight bit][word]: holdings {a in s, c outside s} of the suit with
                                       // two or more board cards, weight = unseen cards of c's rank outside s
    uint32_t YS[4][3];                 // [weight bit][word]: holdings without a card of s (monotone boards)
    uint32_t next_ftask, pd_next, claim_flag;
};

struct Flop4Smem {
    uint16_t flush[kFlushEntries];     // flush[fold(mask)]
    uint16_t pairs[kPairs4 + 3];       // colex pairs i | j << 8 of 47 positions
    uint8_t ppxy[96];                  // rank pair -> x | y << 4
    uint32_t lo_t[14][3], hi_t[14][3]; // rank r -> the rank pairs whose lower / higher rank is r (row 13: empty)
    uint32_t sf3[256];                 // bit m: the 13-bit rank set m has three or more ranks inside some straight window (A2345 .. TJQKA):
                                       // two more cards of the suit could make a straight flush
    Flop4Group G[kGroups];
};

using namespace ms;
__device__ __forceinline__ void group_sync(uint32_t gid) { group_sync_n<kGroupThreads>(gid); }

}  // namespace

template <bool kProf, bool kFastAdd, bool kBits>
__global__ void __launch_bounds__(kThreads4, 1)
hp_flop_treys_v4_kernel(const uint8_t *__restrict__ sit, int64_t n, uint32_t *__restrict__ out, DevTables T,
                        int *__restrict__ other_rows, unsigned long long *__restrict__ prof, uint32_t one, PeerOut peers,
                        unsigned int *__restrict__ work_counter, const uint32_t *__restrict__ order) {
    if (order != nullptr && order[31] == 0) {           // the ordering pass found no flop-shaped row: only tell the other kernels
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            const int other = (order[29] ? 1 : 0) | (order[30] ? 2 : 0);
            if (other) atomicOr(other_rows, other);
        }
        return;
    }
    extern __shared__ __align__(16) uint8_t smem_raw[];
    Flop4Smem &S = *reinterpret_cast<Flop4Smem *>(smem_raw);
    const uint32_t gid = threadIdx.x / kGroupThreads, tid = threadIdx.x % kGroupThreads;
    const uint32_t lane = tid & 31, warp = tid >> 5;
    Flop4Group &G = S.G[gid];
    {
        for (uint32_t m = threadIdx.x; m < kFlushEntries; m += kThreads4) S.flush[fold(m)] = __ldg(T.flush + m);
        for (uint32_t i = threadIdx.x; i < (uint32_t)kPairs4; i += kThreads4) S.pairs[i] = __ldg(T.pairs + i);
        if (threadIdx.x < 96) {
            uint32_t y = 0;
            while (tri(y + 1) <= threadIdx.x) ++y;
            S.ppxy[threadIdx.x] = threadIdx.x < kPP ? (uint8_t)((threadIdx.x - tri(y)) | (y << 4)) : (uint8_t)0xFF;
        }
        if (kBits && threadIdx.x >= 128 && threadIdx.x < 128 + 84) {      // lo_t / hi_t: 2 x 14 rows x 3 words
            const uint32_t k = threadIdx.x - 128, hi_tab = k / 42u, r = (k % 42u) / 3u, w = k % 3u;
            uint32_t m = 0;
            for (uint32_t b = 0; b < 32; ++b) {
                const uint32_t pp = 32u * w + b;
                uint32_t y = 0;
                while (tri(y + 1) <= pp) ++y;
                const uint32_t x = pp - tri(y);
                if (pp < (uint32_t)kPP && (hi_tab ? y : x) == r) m |= 1u << b;
            }
            (hi_tab ? S.hi_t : S.lo_t)[r][w] = m;
        }
        if (kBits) {
            for (uint32_t wd = threadIdx.x >> 5; wd < 256u; wd += kThreads4 / 32) {
                const uint32_t m = 32u * wd + (threadIdx.x & 31u);
                bool three = __popc(m & 0x100Fu) >= 3;                               // A2345
#pragma unroll
                for (int i = 0; i < 9; ++i) three = three || __popc(m & (0x1Fu << i)) >= 3;
                const uint32_t word = __ballot_sync(0xFFFFFFFFu, three);
                if ((threadIdx.x & 31u) == 0) S.sf3[wd] = word;
            }
        }
        if (tid < kProfSlots) G.prof[tid] = 0;
        if (tid == 0) {
            mbar_init((uint32_t)__cvta_generic_to_shared(&G.bar[0]), 1);
            mbar_init((uint32_t)__cvta_generic_to_shared(&G.bar[1]), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    __syncthreads();
    const uint32_t bar0_s = (uint32_t)__cvta_generic_to_shared(&G.bar[0]);
    const uint32_t bar1_s = (uint32_t)__cvta_generic_to_shared(&G.bar[1]);
    uint32_t parity = 0;
    const uint32_t flush_s = (uint32_t)__cvta_generic_to_shared(S.flush);
    const uint32_t t_s = (uint32_t)__cvta_generic_to_shared(G.T);
    int saw_other = 0;
    uint32_t t_prev = 0;
#define HOLDEM_PROF(slot)                                                        \
    if (kProf && tid == 0) {                                                     \
        const uint32_t t_now = (uint32_t)clock64();                              \
        G.prof[slot] += t_now - t_prev;                                          \
        t_prev = t_now;                                                          \
    }

    // Situations are claimed kClaim at a time from a global counter, so that groups which draw expensive boards (monotone
    // flops cost ~3x a rainbow one) simply take fewer: every SM finishes within a few situations of the others.  The next claim
    // is fetched while the current one is being worked on.  (work_counter == nullptr: static striding.)
    const uint32_t n_all_groups = gridDim.x * kGroups, my_group = blockIdx.x * kGroups + gid;
    uint32_t claim_it = 0;
    // a claim = kClaim consecutive work indices; their rows are looked up at once (under the previous situation's task phase)
    auto make_claim = [&](uint32_t slot, uint32_t it) {
        const uint32_t base = work_counter ? atomicAdd(work_counter, (unsigned)kClaim) : (it * n_all_groups + my_group) * kClaim;
        G.next_claim[slot] = base;
#pragma unroll
        for (uint32_t j = 0; j < (uint32_t)kClaim; ++j) {
            const int64_t v = (int64_t)base + j;
            G.claim_rows[slot][j] = v < n ? (order != nullptr ? __ldg(order + kOrderHeader + v) : (uint32_t)v) : 0u;
        }
    };
    if (tid == 0) make_claim(0, 0);
    bool synced = false;           // the previous iteration ended behind barrier (E): its claim and counters are settled
    for (;; ++claim_it) {
    if (!synced) group_sync(gid);
    synced = false;
    const int64_t claim_base = G.next_claim[claim_it & 1];
    if (claim_base >= n) break;
    bool next_claimed = false;
    for (int64_t vidx = claim_base; vidx < n && vidx < claim_base + kClaim; ++vidx) {
        // Work order (kernels_order.cu): monotone flops first, then two-tone, then rainbow, inside a texture the situations
        // in which we hold more cards of one suit first, the rows of other streets last: the heaviest situations start first, and
        // the situation groups of an SM spend most of their time in the same code regions (the kernel is instruction-cache
        // sensitive: +7 % on the mixed workload).  Results are stored by row number as before.
        const int64_t idx = G.claim_rows[claim_it & 1][vidx - claim_base];
        if (kClaim > 1) group_sync(gid);                                                            // (A) (one situation per claim:
                                                                                                    //  the barrier at the loop top serves)
        if (kProf && tid == 0) t_prev = (uint32_t)clock64();
        uint2 w = __ldg(reinterpret_cast<const uint2 *>(sit) + idx);
        uint32_t cb[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) { cb[k] = (w.x >> (8 * k)) & 0xFFu; cb[4 + k] = (w.y >> (8 * k)) & 0xFFu; }
        const bool flop = cb[5] == 0xFFu && cb[6] == 0xFFu;
        if (!flop) { saw_other |= (cb[6] == 0xFFu) ? 1 : 2; continue; }
        bool ok = true;
        uint64_t known = 0, bset = 0;
        uint32_t bsuit = 0, hsuit = 0;   // 4 x 4-bit suit counts (board / hole + board)
        uint32_t rk[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const uint32_t c = cb[k];
            ok = ok && c < 52u;
            const uint32_t cc = c < 52u ? c : 0u;
            known |= card_bit(cc);
            rk[k] = card_rank(cc);
            hsuit += 1u << (4 * card_suit(cc));
            if (k >= 2) {
                bset |= card_bit(cc);
                bsuit += 1u << (4 * card_suit(cc));
            }
        }
        ok = ok && __popcll(known) == 5;
        if (!ok) {
            if (tid == 0) atomicOr(other_rows, 4);          // bit 2: a malformed row was marked
            if (tid < 16) {
                out[idx * 16 + tid] = 0xFFFFFFFFu;
#pragma unroll
                for (int k = 0; k < 8; ++k)      // constant indices: the parameter struct stays in param space
                    if (k < peers.n) peers.p[k][idx * 16 + tid] = 0xFFFFFFFFu;
            }
            continue;
        }
        uint32_t *const gcnt = G.cnt[claim_it & 1];
        if (tid < 12) gcnt[tid] = 0;
        if (tid < 4) { G.bmask[tid] = (uint32_t)(bset >> (13 * tid)) & 0x1FFFu; G.hist_none[tid] = 0; G.Wt0[tid] = 0; G.Ht[tid] = 0; }
        if (tid >= 32 && tid < 96) (&G.Wc0[0][0])[tid - 32] = 0;
        if (tid >= 32 && tid < 128) G.qflush[tid - 32] = 0;
        if (kBits) {
#pragma unroll
            for (int w = 0; w < 3; ++w)
#pragma unroll
                for (int k = 0; k < kBitEntries; k += kGroupThreads) { G.LB[w][k + tid] = 0; G.GB[w][k + tid] = 0; }
        }
        // The set-up below is spread over the group's five warps (a dependent instruction costs ~40 cycles on the busy SM, so the
        // length of the longest chain is what counts): warp 0 sorts the unseen cards, warps 2 and 3 build the per-suit position lists
        // of two suits each, warp 4 lays out the flush task groups, warp 1 fetches the table rows.
        static_assert(kGroupThreads == 160, "the set-up phase assigns work to five warps");
        const uint32_t ltm = (1u << lane) - 1u;
        if (warp == 0) {
            // ---- unseen cards in (rank, suit) order; unseen cards per rank ----------------------------------------------------------
            const uint32_t p0 = path_of_code4(lane), p1 = path_of_code4(lane + 32);
            const bool in0 = !((known >> p0) & 1ull);
            const bool in1 = (lane + 32 < 52) && !((known >> p1) & 1ull);
            const uint32_t m0 = __ballot_sync(0xFFFFFFFFu, in0), m1 = __ballot_sync(0xFFFFFFFFu, in1);
            if (in0) G.code[__popc(m0 & ltm)] = (uint8_t)lane;
            if (in1) G.code[__popc(m0) + __popc(m1 & ltm)] = (uint8_t)(lane + 32);
            {   // rank r owns codes 4r..4r+3
                const uint32_t c4 = 4 * min(lane, 13u);
                const uint32_t cnt_r = lane < 13 ? (c4 < 32 ? __popc((m0 >> c4) & 15u) : __popc((m1 >> (c4 - 32)) & 15u)) : 0u;
                if (lane < 13) G.u[lane] = (uint8_t)cnt_r;
                const uint32_t lo = __reduce_or_sync(0xFFFFFFFFu, lane < 10 ? cnt_r << (3 * lane) : 0u);
                const uint32_t hi = __reduce_or_sync(0xFFFFFFFFu, (lane >= 10 && lane < 13) ? cnt_r << (3 * lane - 30) : 0u);
                if (lane == 0) { G.upack[0] = lo | (hi << 30); G.upack[1] = hi >> 2; }
            }
        } else if (warp == 2 || warp == 3) {
            // ---- per-suit lists of positions in that order: this lane = the cards with codes lane and lane + 32 (suit = lane & 3) ----
            const uint32_t p0 = path_of_code4(lane), p1 = path_of_code4(lane + 32);
            const bool in0 = !((known >> p0) & 1ull);
            const bool in1 = (lane + 32 < 52) && !((known >> p1) & 1ull);
            const uint32_t m0 = __ballot_sync(0xFFFFFFFFu, in0), m1 = __ballot_sync(0xFFFFFFFFu, in1);
            const uint32_t pos0 = __popc(m0 & ltm), pos1 = __popc(m0) + __popc(m1 & ltm);
#pragma unroll
            for (uint32_t h = 0; h < 2; ++h) {
                const uint32_t s = 2u * (warp - 2u) + h;
                const bool a0 = in0 && (lane & 3u) == s, a1 = in1 && (lane & 3u) == s;
                const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, a0), b1 = __ballot_sync(0xFFFFFFFFu, a1);
                const uint32_t k0 = __popc(b0 & ltm), k1 = __popc(b0) + __popc(b1 & ltm);     // cards of s below this one
                if (a0) G.Ls[s][k0] = (uint8_t)pos0;
                else if (in0) G.Lns[s][pos0 - k0] = (uint8_t)pos0;
                if (a1) G.Ls[s][k1] = (uint8_t)pos1;
                else if (in1) G.Lns[s][pos1 - k1] = (uint8_t)pos1;
            }
        } else if (warp == 4) {
            if (lane < 4) {
                const uint32_t rm = ~(uint32_t)(known >> (13 * lane)) & 0x1FFFu;       // ranks of the unseen cards of suit `lane`
                G.smask[lane] = rm;
                G.ns[lane] = (uint8_t)__popc(rm);
            }
            {
                // ---- flush task groups and PD sections: lane = 3 * suit + (2 - qtype), scans over the warp -------------------------
                const uint32_t s = min(lane / 3u, 3u), qtype = 2u - (lane - 3u * (lane / 3u));
                const uint32_t bs = (bsuit >> (4 * s)) & 15u, nsu = __popc(~(uint32_t)(known >> (13 * s)) & 0x1FFFu);
                const uint32_t nboth = choose2u(nsu);
                // step list of the suit: [both cards in s | a | (a, rank of c) | rank pairs without a card of s]
                // (kBits: the undo-only steps are needed by the step-by-step undo half alone and -- their class histograms HA /
                //  hist_none -- by monotone boards)
                const bool short_list = kBits && bs == 2;
                const uint32_t len2 = nboth, len1 = nboth + (short_list ? nsu : 14u * nsu), len0 = len1 + kPP;
                const uint32_t len = bs == 1 ? len2 : (bs == 2 ? len1 : (bs == 3 ? len0 : 0u));
                // PD offset of the suit = lengths of the suits before it (every lane of a suit holds the same len)
                uint32_t pdo = 0;
#pragma unroll
                for (uint32_t q = 0; q < 3; ++q) {
                    const uint32_t lq = __shfl_sync(0xFFFFFFFFu, len, 3 * q);
                    if (q < s) pdo += lq;
                }
                const uint32_t k = bs + qtype;                       // cards of s in board + runout
                const uint32_t need = 5u - k;                        // 2, 1 or 0 cards of s in P
                // Runouts with a card outside s are merged by the RANK of that card (weight = number of such cards): the card matters
                // through its rank alone -- also when the runout completes OUR flush (hs_s + cards of the runout in s >= 5, true for all
                // runouts of a group or none), because that flush is made of our cards, the board's and the runout's card of s
                const uint32_t nq = qtype == 2 ? nboth : (qtype == 1 ? 13u * nsu : (uint32_t)kPP);
                const uint32_t np = need == 2 ? len2 : (need == 1 ? len1 : len0);
                const bool live = lane < 12 && bs != 0 && k >= 3 && nq != 0 && np != 0;
                // (kBits: only the add half of some lanes still walks steps, and those with an add part are few: one task per 32 runouts)
                const uint32_t nseg = kBits ? 1u : (np + kSeg - 1) / kSeg, seglen = live ? (np + nseg - 1) / nseg : 0u;
                const uint32_t ntask = live ? ((nq + 31) / 32) * nseg : 0u;
                // task groups in order of the cards of s still needed (0, 1, 2): the ones with the longest step lists first
                const uint32_t livem = __ballot_sync(0xFFFFFFFFu, live);
                const uint32_t lv0 = __ballot_sync(0xFFFFFFFFu, live && need == 0), lv1 = __ballot_sync(0xFFFFFFFFu, live && need == 1);
                const uint32_t lv2 = livem & ~(lv0 | lv1);
                const uint32_t gi = need == 0 ? __popc(lv0 & ltm)
                                              : (need == 1 ? __popc(lv0) + __popc(lv1 & ltm) : __popc(lv0 | lv1) + __popc(lv2 & ltm));
                if (live) G.g_first[gi + 1] = ntask;
                __syncwarp();
                uint32_t first = (lane >= 1 && lane <= (uint32_t)__popc(livem)) ? G.g_first[lane] : 0u;   // inclusive scan of the counts
#pragma unroll
                for (int d = 1; d < 16; d <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, first, d);
                    if ((int)lane >= d) first += v;
                }
                __syncwarp();
                if (lane <= (uint32_t)__popc(livem) && lane <= (uint32_t)kMaxGroups) G.g_first[lane] = first;
                if (live) {
                    G.g_nq[gi] = nq; G.g_np[gi] = np; G.g_nseg[gi] = nseg; G.g_seglen[gi] = seglen;
                    G.g_pdbase[gi] = pdo;
                    G.g_sq[gi] = s | (qtype << 2);
                }
                if (lane < 12 && qtype == 2) { G.pd_base[s] = pdo; G.pd_len[s] = len; }
                if (lane == 11) {
                    G.n_groups = __popc(livem);
                    G.pd_total = pdo + len;
                    G.next_task = 0;
                    G.next_ftask = 0;
                    G.pd_next = 0;
                    G.claim_flag = 0;
                    G.n_ext = 0;
                }
            }
        } else if (warp == 1) {
            // ---- the situation's table rows, fetched by the TMA engine: T = row b3 of board_t (17 KB, needed by the tasks),
            //      our7nf = row m5 of own_our7, opp5nf = row b3 of board_opp5 (needed at once).  No hash lookups. -----------
            uint32_t ours_now = 0;
            if (lane == 0) {
                uint32_t b0 = rk[2], b1 = rk[3], b2 = rk[4];
                cswap(b0, b1); cswap(b1, b2); cswap(b0, b1);
                const uint32_t b3 = b0 + tri(b1) + (b2 + 2) * (b2 + 1) * b2 / 6;
                uint32_t a0 = rk[0], a1 = rk[1], a2 = b0, a3 = b1, a4 = b2;      // five ranks: insert the hole cards
                cswap(a0, a1);
                cswap(a1, a2); cswap(a2, a3); cswap(a3, a4);                    // the larger hole rank sinks into place
                cswap(a0, a1); cswap(a1, a2); cswap(a2, a3);                    // then the smaller one
                const uint32_t m5 = a0 + tri(a1) + (a2 + 2) * (a2 + 1) * a2 / 6 + (a3 + 3) * (a3 + 2) * (a3 + 1) * a3 / 24 +
                                    (a4 + 4) * (a4 + 3) * (a4 + 2) * (a4 + 1) * a4 / 120;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier reads of the buffers precede the refill
                mbar_expect_tx(bar0_s, 2 * 192);
                bulk_g2s((uint32_t)__cvta_generic_to_shared(G.our7nf), T.own_our7 + (size_t)m5 * 96, 192, bar0_s);
                bulk_g2s((uint32_t)__cvta_generic_to_shared(G.opp5nf), T.board_opp5 + (size_t)b3 * 96, 192, bar0_s);
                mbar_expect_tx(bar1_s, kBoardTRow4 * 2);
                bulk_g2s(t_s, T.board_t + (size_t)b3 * kBoardTRow4, kBoardTRow4 * 2, bar1_s);
                const uint32_t fm = flush_mask(known, 5);
                ours_now = fm ? (uint32_t)S.flush[fold(fm)] : (uint32_t)__ldg(T.own_rank5 + m5);
                G.ours_now = ours_now;
            }
            ours_now = __shfl_sync(0xFFFFFFFFu, ours_now, 0);
            mbar_wait(bar0_s, parity);
            // ---- rank-level flop class per rank pair ------------------------------------------------------------------------
#pragma unroll
            for (uint32_t ps = 0; ps < 3; ++ps) {
                const uint32_t pp = lane + 32u * ps;
                const uint32_t v = G.opp5nf[pp];                     // 0 = the pair cannot exist with this board
                const uint32_t code = v == 0 ? 3u : (ours_now < v ? 0u : (ours_now == v ? 1u : 2u));
                G.clsw[pp] = v == 0 ? 0u : 1u << (10 * code);
                G.clsr[pp] = (uint8_t)code;
            }
        }
        group_sync(gid);                                                                            // (B)
        HOLDEM_PROF(0)
        // the suit in which we can still make a flush (>= 3 of hole + board), and the one the board is monotone in
        const uint32_t f3 = (hsuit + 0x5555u) & 0x8888u;            // nibble >= 3
        const uint32_t fs = f3 ? (uint32_t)(__ffs(f3) - 1) >> 2 : 4u;
        const uint32_t fneed = f3 ? 5u - ((hsuit >> (4 * fs)) & 15u) : 9u;      // cards of fs a runout must bring (<= 0: we hold it)
        const uint32_t hbmask = f3 ? (uint32_t)(known >> (13 * fs)) & 0x1FFFu : 0u;
        const uint32_t m3 = (bsuit + 0x5555u) & 0x8888u;
        const uint32_t ms = m3 ? (uint32_t)(__ffs(m3) - 1) >> 2 : 4u;
        {
            mbar_wait(bar0_s, parity);                                 // (already complete: warp 1 waited before the barrier)
            const uint32_t ours_now = G.ours_now;
            // ---- hand strength: opponent holdings by flop class, from rank pairs (no per-pair table is built) ---------------------
            if (warp < 3) {
                const uint32_t pp = tid;                               // 96 threads, 91 rank pairs
                const uint32_t xy = S.ppxy[min(pp, 90u)], x = xy & 15u, y = xy >> 4;
                const uint32_t ux = G.u[x], uy = G.u[y];
                const uint32_t cnt = pp < (uint32_t)kPP ? (x != y ? ux * uy : (ux ? choose2u(ux) : 0u)) : 0u;
                const uint32_t cl = pp < (uint32_t)kPP ? (uint32_t)G.clsr[min(pp, 90u)] : 3u;
#pragma unroll
                for (uint32_t c = 0; c < 3; ++c) {
                    const uint32_t v = __reduce_add_sync(0xFFFFFFFFu, cl == c ? cnt : 0u);
                    if (lane == 0 && v) atomicAdd(&gcnt[6 + c], v);
                }
                if (kBits) {
                    // ---- the undo half of the flush part as sets of rank pairs (this thread = rank pair pp = (x <= y)) -------------
                    uint32_t s2 = 4;                                                     // the suit with two or more board cards
#pragma unroll
                    for (uint32_t c = 0; c < 3; ++c) {
                        const uint32_t m = __ballot_sync(0xFFFFFFFFu, cl == c);
                        if (lane == 0) G.CL[c][warp] = m;
                    }
#pragma unroll
                    for (uint32_t s = 0; s < 4; ++s) {
                        const uint32_t bs = (bsuit >> (4 * s)) & 15u;
                        if (bs == 0) continue;                                           // (uniform over the group)
                        if (bs >= 2) s2 = s;
                        const uint32_t sm = G.smask[s];
                        const uint32_t m = __ballot_sync(0xFFFFFFFFu, x != y && ((sm >> x) & (sm >> y) & 1u));
                        if (lane == 0) G.BS[s][warp] = m;
                    }
                    if (s2 < 4) {
                        const uint32_t sm = G.smask[s2], ax = (sm >> x) & 1u, ay = (sm >> y) & 1u;
                        const uint32_t bx = ux - ax, by = uy - ay;                       // unseen cards of the rank outside s (<= 3)
                        const uint32_t w1 = ax ? by : 0u;                                // {a = the lower rank's card of s, c of rank y}
                        const uint32_t w2 = (x != y && ay) ? bx : 0u;                    // {a = the higher rank's card of s, c of rank x}
#pragma unroll
                        for (uint32_t k = 0; k < 2; ++k) {
                            const uint32_t m1 = __ballot_sync(0xFFFFFFFFu, (w1 >> k) & 1u);
                            const uint32_t m2 = __ballot_sync(0xFFFFFFFFu, (w2 >> k) & 1u);
                            if (lane == 0) { G.XS[0][k][warp] = m1; G.XS[1][k][warp] = m2; }
                        }
                        if (m3) {
                            const uint32_t wy = x != y ? bx * by : (bx ? choose2u(bx) : 0u);         // no card in s (<= 9)
#pragma unroll
                            for (uint32_t k = 0; k < 4; ++k) {
                                const uint32_t m = __ballot_sync(0xFFFFFFFFu, (wy >> k) & 1u);
                                if (lane == 0) G.YS[k][warp] = m;
                            }
                        }
                    }
                }
            }
            if (ms < 4) {   // monotone board: a holding suited in that suit has flopped a flush -- move it to its true class
                const uint32_t nsu = G.ns[ms], mbmask = G.bmask[ms];
                for (uint32_t t = tid; t < choose2u(nsu); t += kGroupThreads) {
                    const uint32_t pr = S.pairs[t];
                    const uint32_t ra = G.code[G.Ls[ms][pr & 0xFFu]] >> 2, rb = G.code[G.Ls[ms][pr >> 8]] >> 2;
                    const uint32_t opp5 = S.flush[fold(mbmask | (1u << ra) | (1u << rb))];
                    const uint32_t cls = ours_now < opp5 ? 0u : (ours_now == opp5 ? 1u : 2u);
                    atomicAdd(&gcnt[6 + cls], 1u);
                    atomicSub(&gcnt[6 + G.clsr[tri(rb) + ra]], 1u);
                }
            }
            HOLDEM_PROF(1)
            // ---- runouts that give us a flush get their own entry of the generic part (their our7 is a suited-table rank) ---------
            if (fneed <= 2) {
                const uint32_t nfs = G.ns[fs], nboth = choose2u(nfs);
                // A runout's cards outside fs do not change our flush, so the runouts are merged by the RANKS of those cards (entry weight
                // = number of such cards / pairs of cards): C(nfs,2) entries for the runouts inside fs (fneed <= 2), 13 per card a of fs
                // for {a, z outside fs} (fneed <= 1), 91 rank pairs for the runouts without a card of fs (fneed == 0: we hold the flush)
                const uint32_t n1 = nboth + 13u * nfs;
                const uint32_t total = (int)fneed == 2 ? nboth : ((int)fneed == 1 ? n1 : n1 + (uint32_t)kPP);
                if (tid == 0) G.n_ext = total;
                const uint32_t smf = G.smask[fs];
                for (uint32_t t = tid; t < total; t += kGroupThreads) {
                    uint32_t pp, fbits, wgt = 1;
                    if (t < nboth) {
                        const uint32_t pr = S.pairs[t];
                        const uint32_t ri = G.code[G.Ls[fs][pr & 0xFFu]] >> 2, rj = G.code[G.Ls[fs][pr >> 8]] >> 2;
                        fbits = (1u << ri) | (1u << rj);
                        pp = tri(rj) + ri;
                    } else if (t < n1) {
                        const uint32_t t2 = t - nboth, a = t2 / 13u, rz = t2 - 13u * a;
                        const uint32_t ra = G.code[G.Ls[fs][a]] >> 2;
                        pp = tri(max(ra, rz)) + min(ra, rz);
                        fbits = 1u << ra;
                        wgt = G.u[rz] - ((smf >> rz) & 1u);
                    } else {
                        pp = t - n1;
                        const uint32_t xy = S.ppxy[pp], x = xy & 15u, y = xy >> 4;
                        const uint32_t wx = G.u[x] - ((smf >> x) & 1u), wy = G.u[y] - ((smf >> y) & 1u);
                        fbits = 0;
                        wgt = x != y ? wx * wy : (wx ? choose2u(wx) : 0u);
                    }
                    // entry = rank pair | our7 (a suited-table rank) << 8 | weight << 21; slot = index of the runout in its list
                    G.ext[t] = pp | ((uint32_t)S.flush[fold(hbmask | fbits)] << 8) | (wgt << 21);
                    if (wgt) atomicAdd(&G.qflush[pp], wgt);
                }
            }
            HOLDEM_PROF(4)
            // ---- step descriptors of the flush part ---------------------------------------------------------------------------------
            // (32 descriptors per claim: the warps that had nothing to do above start at once, the others join when they get here)
            const uint32_t total = G.pd_total;
            for (;;) {
                uint32_t e = 0;
                if (lane == 0) e = atomicAdd(&G.pd_next, 32u);
                e = __shfl_sync(0xFFFFFFFFu, e, 0) + lane;
                if (e - lane >= total) break;
                if (e >= total) continue;
                uint32_t s = 0;
#pragma unroll
                for (uint32_t q = 1; q < 4; ++q)
                    if (G.pd_len[q] && e >= G.pd_base[q]) s = q;
                uint32_t t = e - G.pd_base[s];
                const uint32_t nsu = G.ns[s], nboth = choose2u(nsu), sm = G.smask[s];
                if (t < nboth) {                                   // both cards of the holding in s
                    const uint32_t pr = S.pairs[t];
                    const uint32_t pa = G.Ls[s][pr & 0xFFu], pb = G.Ls[s][pr >> 8];
                    const uint32_t ra = G.code[pa] >> 2, rb = G.code[pb] >> 2;
                    const uint32_t pp = tri(rb) + ra, bits = (1u << ra) | (1u << rb);
                    uint32_t cls = G.clsr[pp];
                    if (s == ms) {                                 // monotone board: this holding has flopped a flush
                        const uint32_t opp5 = S.flush[fold(G.bmask[s] | bits)];
                        cls = ours_now < opp5 ? 0u : (ours_now == opp5 ? 1u : 2u);
                    }
                    G.PD[e] = make_uint4((bits << 17) | (2u * fold(bits)), 1u << (10 * cls), G.clsw[pp], pp * (kTPitch * 2));
                    atomicAdd(&G.Wc0[s][pr & 0xFFu], 1u << (10 * cls));
                    atomicAdd(&G.Wc0[s][pr >> 8], 1u << (10 * cls));
                    {   // the suit's total: one atomic per warp when the lanes in this branch all work on one suit (the usual case)
                        const uint32_t act = __activemask();
                        int same;
                        __match_all_sync(act, s, &same);
                        if (same) {
                            const uint32_t sum = __reduce_add_sync(act, 1u << (10 * cls));
                            if (lane == (uint32_t)__ffs(act) - 1u) atomicAdd(&G.Wt0[s], sum);
                        } else atomicAdd(&G.Wt0[s], 1u << (10 * cls));
                    }
                } else if (t < nboth + nsu) {                      // one card a in s, any card c outside s: the flop class of {a, c}
                    const uint32_t ra = G.code[G.Ls[s][t - nboth]] >> 2;     // depends on ranks alone (four of the suit at most)
                    uint32_t hist = 0;
                    for (uint32_t y = 0; y < 13; ++y)
                        hist += (G.u[y] - ((sm >> y) & 1u)) * G.clsw[tri(max(ra, y)) + min(ra, y)];
                    G.PD[e] = make_uint4(((2u << ra) << 16) | (2u * fold(1u << ra)), hist, 0u, 0u);
                    G.Hs[s][t - nboth] = hist;
                    atomicAdd(&G.Ht[s], hist);
                } else if (t < nboth + 14u * nsu) {                // undo step for {a, any c of rank y outside s}
                    t -= nboth + nsu;
                    const uint32_t ai = t / 13u, y = t - 13u * ai;
                    const uint32_t ra = G.code[G.Ls[s][ai]] >> 2;
                    const uint32_t pp = tri(max(ra, y)) + min(ra, y);
                    const uint32_t hist = (G.u[y] - ((sm >> y) & 1u)) * G.clsw[pp];
                    G.PD[e] = make_uint4(((2u << ra) << 16) | (2u * fold(1u << ra)), 0u, hist, pp * (kTPitch * 2));
                    if (s == ms) G.HA[ai * 13 + y] = hist;
                } else {                                           // undo step for a rank pair, neither card in s
                    const uint32_t pp = t - (nboth + 14u * nsu);
                    const uint32_t xy = S.ppxy[pp], x = xy & 15u, y = xy >> 4;
                    const uint32_t wx = G.u[x] - ((sm >> x) & 1u), wy = G.u[y] - ((sm >> y) & 1u);
                    const uint32_t wgt = x != y ? wx * wy : (wx ? choose2u(wx) : 0u);
                    G.PD[e] = make_uint4(0u, 0u, wgt * G.clsw[pp], pp * (kTPitch * 2));
                    if (wgt) atomicAdd(&G.hist_none[G.clsr[pp]], wgt);
                }
            }
        }
        HOLDEM_PROF(5)
        group_sync(gid);                                                                            // (D)
        // The next claim (a global atomic, then the row number through the work order: two dependent global accesses) is made by the
        // first warp that runs out of flush tasks, in the time it would otherwise wait at barrier (E) -- not at the start of the task
        // phase, where the claiming warp arrived late at the barrier between the generic and the flush tasks.
        const bool claim_here = !next_claimed;
        next_claimed = true;
        mbar_wait(bar1_s, parity);                                     // T has landed
        parity ^= 1u;
        HOLDEM_PROF(2)

        uint32_t accL[3] = {0, 0, 0}, accG[3] = {0, 0, 0};   // per class: our7 < R / our7 > R (net of the undo steps, mod 2^32)
        const uint32_t n_groups = G.n_groups;
        const uint32_t n_flush_tasks = G.g_first[n_groups];
        const uint32_t n_entries = kPP + G.n_ext;
        // generic tasks first: the large parts (y >= 7) of every chunk of 32 entries, then the small parts (y <= 6)
        const uint32_t n_gen_tasks = 2u * ((n_entries + 31) / 32);
        const uint32_t n_tasks = n_gen_tasks + n_flush_tasks;
        // kBits: the flush tasks read the compare bits that the generic tasks leave behind, so the group meets in between
        uint32_t phase = 0;
        for (;;) {
            uint32_t task = 0;
            if (lane == 0) task = atomicAdd(kBits && phase ? &G.next_ftask : &G.next_task, 1u);
            task = __shfl_sync(0xFFFFFFFFu, task, 0);
            if (kBits) {
                if (phase == 0) {
                    if (task >= n_gen_tasks) { phase = 1; group_sync(gid); HOLDEM_PROF(6) continue; }
                } else if (task >= n_flush_tasks) {
                    if (claim_here && lane == 0 && atomicExch(&G.claim_flag, 1u) == 0u) make_claim((claim_it + 1) & 1, claim_it + 1);
                    break;
                }
            } else {
                if (task >= n_tasks) {
                    if (claim_here && lane == 0 && atomicExch(&G.claim_flag, 1u) == 0u) make_claim((claim_it + 1) & 1, claim_it + 1);
                    break;
                }
                phase = task >= n_gen_tasks ? 1u : 0u;
                if (phase) task -= n_gen_tasks;
            }
            if (phase) {
                // ================= flush task: 32 runouts Q x a segment of steps =======================================
                // the task group: the last one whose first task is <= task (lanes 1.. compare one boundary each)
                const uint32_t g = __popc(__ballot_sync(0xFFFFFFFFu, lane >= 1 && lane < n_groups && task >= G.g_first[min(lane, (uint32_t)kMaxGroups)]));
                const uint32_t local = task - G.g_first[g];
                const uint32_t nseg = G.g_nseg[g], seglen = G.g_seglen[g], np = G.g_np[g], nq = G.g_nq[g];
                const uint32_t chunk = nseg == 1 ? local : local / nseg, seg = local - chunk * nseg;
                const uint32_t s = G.g_sq[g] & 3u, qtype = (G.g_sq[g] >> 2) & 3u;
                const uint32_t p0 = seg * seglen, p1 = min(np, p0 + seglen);
                const uint32_t qi = chunk * 32 + lane;
                const uint32_t nsu = G.ns[s], sm = G.smask[s];
                // the lane's runout: our rank, its cards of s (rank bits), folded flush index of board + runout in s, its column
                // of T, its multiplicity; rz = the rank of its card outside s for the fix-up below
                uint32_t our = 0, fqf = 0, qb2 = 0, rz = 0, qcol = 0, wl = 0, rq = 99, easy = 1;
                uint32_t ia = 0xFFu, ib = 0xFFu;       // the runout's cards of s as indices into Ls[s] (fast path below)
                if (qi < nq) {
                    uint32_t qbits, ppq;
                    const bool ours = ((hsuit >> (4 * s)) & 15u) + qtype >= 5u;        // the group's runouts complete our flush (s == fs)
                    if (qtype == 1) {                          // (card q of s, rank of the other card)
                        const uint32_t a = qi / 13u;
                        ia = a;
                        rz = qi - 13u * a;
                        rq = G.code[G.Ls[s][a]] >> 2;
                        qbits = 1u << rq;
                        ppq = tri(max(rq, rz)) + min(rq, rz);
                        wl = G.u[rz] - ((sm >> rz) & 1u);
                        our = ours ? (uint32_t)S.flush[fold(hbmask | qbits)] : (uint32_t)G.our7nf[ppq];
                    } else if (qtype == 0) {                   // rank pair, neither card in s
                        const uint32_t xy = S.ppxy[qi], x = xy & 15u, y = xy >> 4;
                        const uint32_t wx = G.u[x] - ((sm >> x) & 1u), wy = G.u[y] - ((sm >> y) & 1u);
                        qbits = 0;
                        ppq = qi;
                        wl = x != y ? wx * wy : (wx ? choose2u(wx) : 0u);
                        our = ours ? (uint32_t)S.flush[fold(hbmask)] : (uint32_t)G.our7nf[ppq];
                    } else {                                   // both cards in s
                        const uint32_t pr = S.pairs[qi];
                        ia = pr & 0xFFu; ib = pr >> 8;
                        const uint32_t ri = G.code[G.Ls[s][ia]] >> 2, rj = G.code[G.Ls[s][ib]] >> 2;
                        qbits = (1u << ri) | (1u << rj);
                        ppq = tri(rj) + ri;
                        // our rank on board + this runout: a suited-table rank when it completes our flush (in s, or the one we hold)
                        const uint32_t fb = s == fs ? qbits : 0u;
                        our = (int)__popc(fb) >= (int)fneed ? (uint32_t)S.flush[fold(hbmask | fb)] : (uint32_t)G.our7nf[ppq];
                        wl = 1;
                    }
                    qb2 = qbits << 17;
                    const uint32_t mbq = G.bmask[s] | qbits;
                    fqf = 2u * fold(mbq);
                    qcol = 2u * ppq;
                    // the add parts of the lane in closed form?  Every flush step ranks the opponent by a real flush (323..1599, or a
                    // straight flush 1..10).  our > 1599: we lose every one of them; our <= 322 (quads / full house, never a flush on
                    // this path unless the runout completes ours): we win every one unless a straight flush is possible at all.
                    easy = our > 1599u ? 1u : ((kBits && our > 10u && our <= 322u && !((S.sf3[mbq >> 5] >> (mbq & 31u)) & 1u)) ? 2u : 0u);
                }
                const uint32_t tq_s = t_s + qcol;
                const uint4 *pd = G.PD + G.g_pdbase[g] + p0;
                uint32_t aL = 0, aG = 0, sL = 0, sG = 0;
                const uint32_t need = 5u - (((bsuit >> (4 * s)) & 15u) + qtype);
                // Runouts that complete OUR flush (all of the task group or none) have their own entries in the generic part and a
                // suited rank: they keep the step-by-step undo half.  Everywhere else the lane's runout shares our7 and the column of T
                // with merged runout ppq of the generic part, whose compare outcomes LB / GB are already there: the undo half of the whole
                // step list is a weighted population count.  Step weights by rank pair: both cards in s -> 1 (BC), {a in s, c outside s}
                // -> unseen cards of c's rank outside s (XC, by which of the two ranks is a's), no card in s -> YC; a step is valid when
                // its cards of s avoid the runout's (lo_t / hi_t rows of the runout's ranks in s).
                const uint32_t hs_s = (hsuit >> (4 * s)) & 15u;
                const bool ext_lanes = hs_s + qtype >= 5u;                   // every runout of the group completes our flush (s == fs)
                if (kBits) {
                    const uint32_t split = min(p1, max(p0, choose2u(nsu) + nsu));
                    const bool fast = kFastAdd && __all_sync(0xFFFFFFFFu, easy != 0u);
                    if (seg == 0) {
                        const uint32_t qbits = qb2 >> 17, rest = qbits & (qbits - 1u);
                        // the lane's entry of the generic part: merged runout ppq, or -- in the order the our-flush entries were written --
                        // [pairs inside s | 13 per card of s | 91 rank pairs without a card of s]
                        const uint32_t ext_first = qtype == 2 ? 0u : choose2u(nsu) + (qtype == 1 ? 0u : 13u * nsu);
                        const uint32_t ppq = ext_lanes ? (qi < nq ? (uint32_t)kPP + ext_first + qi : 0u) : qcol >> 1;
                        const uint32_t r1 = qbits ? (uint32_t)__ffs(qbits) - 1u : 13u, r2 = rest ? (uint32_t)__ffs(rest) - 1u : 13u;
#pragma unroll 1
                        for (uint32_t w = 0; w < 3; ++w) {
                            const uint32_t lb = G.LB[w][ppq], gb = G.GB[w][ppq];
                            const uint32_t lo = ~(S.lo_t[r1][w] | S.lo_t[r2][w]), hi = ~(S.hi_t[r1][w] | S.hi_t[r2][w]);
                            const uint32_t cl0 = G.CL[0][w], cl1 = G.CL[1][w], cl2 = G.CL[2][w];
                            // (rank pairs that cannot exist with this board are in no class: their weights never count)
                            auto tally = [&](uint32_t ml, uint32_t mg, uint32_t k) {
                                sL += ((uint32_t)__popc(ml & cl0) << k) + ((uint32_t)__popc(ml & cl1) << (10 + k)) +
                                      ((uint32_t)__popc(ml & cl2) << (20 + k));
                                sG += ((uint32_t)__popc(mg & cl0) << k) + ((uint32_t)__popc(mg & cl1) << (10 + k)) +
                                      ((uint32_t)__popc(mg & cl2) << (20 + k));
                            };
                            const uint32_t both = G.BS[s][w] & lo & hi;
                            tally(lb & both, gb & both, 0);
                            if (need <= 1) {
#pragma unroll
                                for (uint32_t k = 0; k < 2; ++k) {
                                    const uint32_t m1 = G.XS[0][k][w] & lo, m2 = G.XS[1][k][w] & hi;
                                    tally(lb & m1, gb & m1, k);
                                    tally(lb & m2, gb & m2, k);
                                }
                                if (need == 0) {
#pragma unroll
                                    for (uint32_t k = 0; k < 4; ++k) {
                                        const uint32_t m = G.YS[k][w];
                                        tally(lb & m, gb & m, k);
                                    }
                                }
                            }
                        }
                    }
                    if (fast) {                                              // add half in closed form (see below)
                        if (seg == 0 && qi < nq) {
                            const bool with_one = np > choose2u(nsu);
                            uint32_t add = G.Wt0[s] + (with_one ? G.Ht[s] : 0u);
                            if (ia != 0xFFu) add -= G.Wc0[s][ia] + (with_one ? G.Hs[s][ia] : 0u);
                            if (ib != 0xFFu) add -= G.Wc0[s][ib] + (with_one ? G.Hs[s][ib] : 0u) - G.PD[G.g_pdbase[g] + qi].y;
                            if (easy == 1u) aG = add;                        // our > suited rank for every one of them
                            else aL = add;                                   // quads / full house against flushes
                        }
                    } else {                                                 // add half step by step: suited-table gather, two compares
#pragma unroll kF4Unroll
                        for (uint32_t t = p0; t < split; ++t, ++pd) {
                            const uint4 d = *pd;
                            undo_step(aL, aG, our, lds_u16(flush_s + ((d.x & 0xFFFFu) ^ fqf)), d.x, qb2, d.y, one);
                        }
                    }
                } else {
                // steps [0, nboth + n_s) have an add part (holdings with both cards in s, or one card a in s); the ones behind them
                // are rank-level undo steps: no suited-table gather, half the compares
                const uint32_t split = min(p1, max(p0, choose2u(nsu) + nsu));
                // Fast path.  Every step with an add part ranks the opponent by a real flush (rank <= 1599).  When none of the warp's
                // runouts gives US a flush, full house or quads (our rank > 1599: the common case), every valid add part is a loss for us
                // whatever the exact suited rank, so the add parts of the whole step list sum to a closed form per lane: the class
                // histogram of all such holdings minus those that contain one of the runout's cards of s (inclusion-exclusion over
                // Wc0 / Hs).  The steps then only carry their undo half: no suited-table gather, half the compares.
                const bool fast = kFastAdd && __all_sync(0xFFFFFFFFu, our > 1599u || qi >= nq);
                if (fast) {
#pragma unroll kF4Unroll
                    for (uint32_t t = p0; t < min(split, max(p0, choose2u(nsu))); ++t, ++pd) {   // (the "card a in s" steps have no undo half)
                        const uint4 d = *pd;
                        undo_step(sL, sG, our, lds_u16(tq_s + d.w), d.x, qb2, d.z, one);
                    }
                    pd = G.PD + G.g_pdbase[g] + split;
                    if (seg == 0 && qi < nq) {
                        const bool with_one = np > choose2u(nsu);            // the list also holds the "card a in s" steps
                        uint32_t add = G.Wt0[s] + (with_one ? G.Ht[s] : 0u);
                        if (ia != 0xFFu) add -= G.Wc0[s][ia] + (with_one ? G.Hs[s][ia] : 0u);
                        if (ib != 0xFFu) add -= G.Wc0[s][ib] + (with_one ? G.Hs[s][ib] : 0u) - G.PD[G.g_pdbase[g] + qi].y;
                        aG = add;                                            // our > suited rank for every one of them
                    }
                } else {
#pragma unroll kF4Unroll
                    for (uint32_t t = p0; t < split; ++t, ++pd) {
                        const uint4 d = *pd;
                        // fold() is XOR-linear and the step's cards of s are disjoint from board + runout when the step is valid
                        const uint32_t rf = lds_u16(flush_s + ((d.x & 0xFFFFu) ^ fqf));
                        const uint32_t rn = lds_u16(tq_s + d.w);
                        flush_step(aL, aG, sL, sG, our, rf, rn, d.x, qb2, d.y, d.z, one);
                    }
                }
#pragma unroll 2
                for (uint32_t t = split; t < p1; ++t, ++pd) {
                    const uint4 d = *pd;
                    undo_step(sL, sG, our, lds_u16(tq_s + d.w), d.x, qb2, d.z, one);
                }
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) {   // wl = 0 on lanes past the end of the runout list
                    accL[c] += (((aL >> (10 * c)) & 0x3FFu) - ((sL >> (10 * c)) & 0x3FFu)) * wl;
                    accG[c] += (((aG >> (10 * c)) & 0x3FFu) - ((sG >> (10 * c)) & 0x3FFu)) * wl;
                }
                if (seg == 0 && need == 0 && qi < nq) {
                    // holdings without a card of s: the board + runout flush itself, one compare for all of them
                    const uint32_t rf = lds_u16(flush_s + fqf);
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        if (our < rf) accL[c] += G.hist_none[c];
                        if (our > rf) accG[c] += G.hist_none[c];
                    }
                }
                if (seg == 0 && need == 1 && qtype == 1) {
                    // every card z of rank rz outside s was counted among the c of {a, c} in its own runout: take the pairs
                    // {a, z} out again -- summed over z they are the class histogram HA[a][rz] and wl undo steps of (ra, rz)
                    for (uint32_t ai = 0; ai < nsu; ++ai) {
                        const uint32_t ra = G.code[G.Ls[s][ai]] >> 2;
                        if (wl != 0 && ra != rq) {
                            const uint32_t rf = lds_u16(flush_s + (fqf ^ (2u * fold(1u << ra))));
                            const uint32_t pp = tri(max(ra, rz)) + min(ra, rz);
                            const uint32_t rn = lds_u16(tq_s + pp * (kTPitch * 2));
                            const uint32_t ha = G.HA[ai * 13 + rz], cw = wl * G.clsw[pp];
#pragma unroll
                            for (uint32_t c = 0; c < 3; ++c) {
                                accL[c] += (our < rn ? (cw >> (10 * c)) & 0x3FFu : 0u) - (our < rf ? (ha >> (10 * c)) & 0x3FFu : 0u);
                                accG[c] += (our > rn ? (cw >> (10 * c)) & 0x3FFu : 0u) - (our > rf ? (ha >> (10 * c)) & 0x3FFu : 0u);
                            }
                        }
                    }
                }
            } else {
                // ================= generic task: 32 merged runouts, the warp walks the rank pairs of P class by class =====
                // task -> 32 entries x one part of the rank pairs, as a run of slices of two to seven values of y.  The two parts of a
                // chunk are UNEQUAL on purpose (y >= 7: 63 rank pairs, y <= 6: 28) and the large parts come first in the queue: with
                // three chunks (no runout completes our flush: 63 % of the flops) the group's five warps take 3 large + 2 small parts
                // and the first to finish takes the last small one, so the phase lasts 63 steps instead of 2 x 45; with five chunks
                // every warp gets a large and a small part as before.  (Cutting the three chunks into 3 x 5 single-slice tasks was
                // measured and is slower, 52.1 -> 48.9 M situations/s: the per-task set-up chain costs more than the imbalance.)
                const uint32_t n_chunks = n_gen_tasks >> 1;
                const bool small_part = task >= n_chunks;
                const uint32_t e = (small_part ? task - n_chunks : task) * 32 + lane;
                const uint32_t sl = small_part ? 0u : 1u, sl_end = small_part ? 1u : 4u;
                uint32_t qp = 0, val = 0, mult = 0;
                if (e < (uint32_t)kPP) {
                    qp = e;
                    val = G.our7nf[e];
                } else if (e < n_entries) {
                    const uint32_t en = G.ext[e - kPP];
                    qp = en & 0xFFu;
                    val = (en >> 8) & 0x1FFFu;
                    mult = en >> 21;
                }
                const uint32_t qxy = S.ppxy[qp], q1 = qxy & 15u, q2 = qxy >> 4;
                if (e < (uint32_t)kPP) {
                    const uint32_t u1 = G.u[q1], u2 = G.u[q2];
                    mult = (q1 != q2 ? u1 * u2 : (u1 ? choose2u(u1) : 0u)) - G.qflush[qp];
                }
                // unseen cards per rank once the runout's two cards are gone, 3 bits per rank
                const uint64_t m64 = (((uint64_t)G.upack[1] << 32) | G.upack[0]) - (1ull << (3 * q1)) - (1ull << (3 * q2));
                const uint32_t tq_s = t_s + 2u * qp;
                // the rank pairs of P, fully unrolled inside a slice (constant shifts and offsets); unweighted counts n_Q(pp) go into
                // packed 10-bit class fields (they sum to C(45,2) = 990 per runout), the runout multiplicity is applied afterwards
                uint32_t lp = 0, gp = 0;
                // kBits: the compare outcomes of a slice, bit pp of word pp / 32 (a slice touches words W0 and W0 + 1 at most), are OR-ed
                // into the entry's zeroed sets at the end of the slice
#define HOLDEM_GEN_SLICE(Y0, Y1, W0)                                                                                                \
    {                                                                                                                               \
        uint32_t lb0 = 0, lb1 = 0, gb0 = 0, gb1 = 0;                                                                                \
        _Pragma("unroll") for (int y = Y0; y <= Y1; ++y) {                                                                          \
            const uint32_t my = (uint32_t)(m64 >> (3 * y)) & 7u;                                                                    \
            _Pragma("unroll") for (int x = 0; x <= y; ++x) {                                                                        \
                const int pp = y * (y + 1) / 2 + x;                                                                                 \
                const uint32_t nn = x == y ? (my * (my - 1u)) >> 1 : ((uint32_t)(m64 >> (3 * x)) & 7u) * my;                        \
                const uint32_t R = lds_u16(tq_s + pp * (kTPitch * 2));                                                              \
                if (!kBits) cmp_add(lp, gp, val, R, nn, G.clsw[pp]);                                                                \
                else if ((pp >> 5) == W0) cmp_add_bits(lp, gp, lb0, gb0, val, R, nn, G.clsw[pp], 1u << (pp & 31));                  \
                else cmp_add_bits(lp, gp, lb1, gb1, val, R, nn, G.clsw[pp], 1u << (pp & 31));                                       \
            }                                                                                                                       \
        }                                                                                                                           \
        if (kBits && e < (uint32_t)kBitEntries) {                                                                                   \
            atomicOr(&G.LB[W0][e], lb0); atomicOr(&G.GB[W0][e], gb0);                                                               \
            if ((Y1 * (Y1 + 3) / 2) >> 5 != W0) { atomicOr(&G.LB[W0 + 1][e], lb1); atomicOr(&G.GB[W0 + 1][e], gb1); }               \
        }                                                                                                                           \
    }
                if (sl == 0) HOLDEM_GEN_SLICE(0, 6, 0)                               // rank pairs  0..27
                if (sl <= 1 && sl_end > 1) HOLDEM_GEN_SLICE(7, 8, 0)                 //            28..44
                if (sl <= 2 && sl_end > 2) HOLDEM_GEN_SLICE(9, 10, 1)                //            45..65
                if (sl_end > 3) HOLDEM_GEN_SLICE(11, 12, 2)                          //            66..90
#undef HOLDEM_GEN_SLICE
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    accL[c] += ((lp >> (10 * c)) & 0x3FFu) * mult;
                    accG[c] += ((gp >> (10 * c)) & 0x3FFu) * mult;
                }
            }
        }
        // ---- reduce over the warp, then into the group's counters ------------------------------------------------------------------
#pragma unroll
        for (uint32_t c = 0; c < 3; ++c) {
            const uint32_t l = __reduce_add_sync(0xFFFFFFFFu, accL[c]);
            const uint32_t g2 = __reduce_add_sync(0xFFFFFFFFu, accG[c]);
            if (lane == 0) {
                if (l) atomicAdd(&gcnt[c], l);
                if (g2) atomicAdd(&gcnt[3 + c], g2);
            }
        }
        group_sync(gid);                                                                            // (E)
        synced = kClaim == 1;
        HOLDEM_PROF(3)
        if (tid < 16) {
            uint32_t v;
            if (tid < 3) v = gcnt[6 + tid];
            else if (tid < 12) {
                const uint32_t cls = (tid - 3) / 3, o = (tid - 3) % 3;
                const uint32_t lt = gcnt[cls], gt = gcnt[3 + cls], all = 990u * gcnt[6 + cls];
                v = o == 0 ? lt : (o == 2 ? gt : all - lt - gt);
            } else if (tid < 15) v = gcnt[6 + tid - 12];
            else v = 3u;
            out[idx * 16 + tid] = v;
#pragma unroll
            for (int k = 0; k < 8; ++k)          // other GPUs' result buffers (NVLink stores)
                if (k < peers.n) peers.p[k][idx * 16 + tid] = v;
        }
    }
    if (!next_claimed && tid == 0) make_claim((claim_it + 1) & 1, claim_it + 1);   // every row of the claim belonged to another kernel
    }
#undef HOLDEM_PROF
    if (saw_other && tid == 0) atomicOr(other_rows, saw_other);
    if (kProf && tid == 0)
        for (int k = 0; k < kProfSlots; ++k) atomicAdd(prof + k, (unsigned long long)G.prof[k]);
}

cudaError_t launch_hp_flop_treys_v4(const DevTables &T, const uint8_t *sit, int64_t n, uint32_t *out, int sms,
                                    int *other_rows_flag, unsigned long long *prof, const PeerOut &peers, cudaStream_t st,
                                    bool closed_form_add, uint32_t *order_scratch, bool bit_parallel_undo) {
    // other_rows_flag points at a 16-byte slot: [0] the flag word, [1] the work counter of this kernel, [2] of the turn kernel
    if (n >= ((int64_t)1 << 31)) return cudaErrorInvalidValue;
    if (n == 0) return cudaSuccess;
    const size_t smem = sizeof(Flop4Smem);
    // the phase clocks are compiled out of the production instances: the kernel is sensitive to its code size
    auto kernel = prof ? hp_flop_treys_v4_kernel<true, true, true>
                       : (!bit_parallel_undo ? hp_flop_treys_v4_kernel<false, true, false>
                          : (closed_form_add ? hp_flop_treys_v4_kernel<false, true, true> : hp_flop_treys_v4_kernel<false, false, true>));
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int64_t blocks = sms;
    if (blocks * kGroups > n) blocks = (n + kGroups - 1) / kGroups;
    e = cudaMemsetAsync(other_rows_flag, 0, 4 * sizeof(int), st);
    if (e != cudaSuccess) return e;
    count_launch();
    kernel<<<(int)blocks, kThreads4, smem, st>>>(sit, n, out, T, other_rows_flag, prof, 1u, peers,
                                                 reinterpret_cast<unsigned int *>(other_rows_flag) + 1, order_scratch);
    return cudaGetLastError();
}

}  // namespace holdem
