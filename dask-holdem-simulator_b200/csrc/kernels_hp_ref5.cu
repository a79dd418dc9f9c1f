// kernels_hp_ref5.cu -- hand strength + hand potential in HOLDEM_MODE_REFERENCE, rank-multiset formulation ("v5"), and the
// rank-pair hand-strength kernel of the same mode.  Flop / turn / river boards.  sm_100a only.
//
// What is counted -- the reference's loops exactly as written (HandCalculations.py:178-226 with the 9-category ranker of
// :45-82, higher = better):
//     HP[cls(P)][cmp(our(Q), opp(P,Q))] += 1   for EVERY opponent holding P (pairs of the m = 47/46/45 cards unseen by us,
//                                               :148-150) and EVERY runout Q (pairs of deck \ board, :152-154: 1176/1128/1081,
//                                               they may repeat our hole cards or the opponent's own cards)
// with our(Q) = category of the card MULTISET hole + board + Q, opp(P,Q) = category of board + Q + P, cls(P) from the
// categories on the current board.  Unlike the canonical enumeration (kernels_hp_flop4.cu) P and Q are independent: the
// sum runs over the full Cartesian product, which is what makes this formulation short.
//
// How (bit-identical counts; the identity is restated on the CPU in scripts/proto_ref_multiset.py and checked against the
// literal oracle by tests/test_multiset_identity.py; on the GPU this kernel is compared with the plain enumeration, with the
// earlier 4-set kernel and with the golden vectors of the unmodified reference):
//   The literal category of a multiset is  NF = the category its RANK counts alone give (no flush tests)  unless a suit holds
//   five or more of its cards (repeats counted), in which case it is 8 when that suit's cards are pairwise distinct and contain
//   a cyclic five-run (:24-37, :61-65), else NF if NF is 7, 6 or 4 (quads / full house / straight are tested before the flush,
//   :68-72), else 5.  So
//     sum over (P,Q)  =  GENERIC part: sum over (rank pair q of Q, our category o) of  M(q,o) x  sum over rank pairs pp of
//                            N(pp) [clsr(pp)] x cmp(o, T[q][pp])
//                        with T[q][pp] = NF(board + q + pp), N(pp) = number of holdings with those ranks (a product of
//                        per-rank counts, independent of Q), M(q,o) = number of runouts with rank pair q on which we hold
//                        category o: ~150 (q,o) entries x 91 compare steps per situation;
//                     +  FLUSH part: for every suit s with b board cards and every (P,Q) with b + |Q in s| + |P in s| >= 5
//                        (a card shared by P and Q counts twice, as in the reference's Counter):
//                            + [cls(P)]  x cmp(our(Q), 8 if no shared card of s and cyclic run in s, else keep(T[q][pp]))
//                            - [clsr(pp)] x cmp(our(Q), T[q][pp] nf)                               (undo the generic term)
//                        Cards outside s matter through their RANK only, on both sides: holdings with a card outside s are
//                        merged into weighted steps, runouts with a card outside s into weighted lanes (counted per
//                        (cards in s, rank of the rest, our category) while our categories are computed).
//                        ~7 K items on a rainbow flop, ~19 K two-tone, ~40 K monotone instead of 1.27 M evaluations.
//
// One CTA of 256 threads per situation (claimed from a global counter: suited boards cost far more than rainbow ones).  Per
// situation: the unseen cards in (rank, suit) order and per-suit position lists; T[91][91] (two nibbles per entry: NF and keep(NF))
// = the board's 8,384-byte row of a device-resident table built at holdem_init, fetched with cp.async.bulk onto an mbarrier
// under the set-up phases; our category on every runout and the class of every holding by the
// arithmetic ranker of device_common.cuh (multiset-exact); then the generic entries and a dynamic queue of flush tasks
// (32 runouts of one (suit, cards-in-suit) class x a segment of the suit's holding descriptors; a step is one broadcast
// descriptor load, one table byte, a bit test and four compare / predicated multiply-add pairs on packed 10-bit class counters).
#include "device_common.cuh"
#include "launch.h"
#include "multiset_common.cuh"

namespace holdem {

namespace {

using namespace ms;

#ifndef HOLDEM_R5_CTAS
#define HOLDEM_R5_CTAS 5
#endif
constexpr int kR5Threads = 256;
constexpr int kTPitch5 = 92;          // bytes per row of T
constexpr int kPP5 = 91;
constexpr int kMaxQ5 = 1176;          // C(49,2)
constexpr int kMaxP5 = 1081;          // C(47,2)
constexpr int kPDCap5 = 512;          // holding descriptors of all suits: C(n,2) [+ 13 n [+ 91]] per suit by its board cards; worst case 464 (river board 3 + 2)
constexpr int kSeg5 = 96;             // steps per flush task: a step adds at most 9 to a packed 10-bit counter (1023)
constexpr int kTRow5 = 8384;          // bytes per board of the global T tables (91 x 92 rounded up to a multiple of 16)
constexpr int kMaxTaskGroups = 12;    // (suit, cards of Q in suit) classes

struct Ref5Smem {
    uint4 PD[kPDCap5];                 // x = rank bits of the holding's cards in s, y = 1 << 10 cls(P), z = 1 << 10 clsr(pp), w = pp
    uint32_t Wt32[96];                 // per rank pair: N(pp) << 10 clsr(pp)  (N <= 16)
    unsigned long long bar;            // mbarrier of the T row copy
    uint32_t M[kPP5 * 9 + 1];          // runouts by (rank pair, our category)
    uint16_t elist[kPP5 * 9 + 1];      // nonzero entries of M, compacted
    uint32_t cyc[256];                 // bit m: the 13-bit rank mask m holds a cyclic five-run (static)
    uint32_t cnt[12];                  // behind[3], ahead[3] (by class), class totals[3]
    uint32_t bmask[4];                 // board rank mask per suit
    uint32_t g_first[kMaxTaskGroups + 1];
    uint32_t g_nq[kMaxTaskGroups], g_np[kMaxTaskGroups], g_nseg[kMaxTaskGroups], g_pdbase[kMaxTaskGroups], g_sq[kMaxTaskGroups];
    uint32_t pd_base[4], pd_len[4];
    uint32_t h_np[kMaxTaskGroups], h_pdo[kMaxTaskGroups];   // per (suit, qtype) class before the live ones are compacted
    uint32_t n_groups, n_entries, next_task, claim;
    uint32_t nL1[2], nL0;              // merged runout lanes per list
    uint32_t A1[2][192];               // runouts with ONE card in a suit holding >= 2 board cards, counted by (rank of that card, rank
                                       // of the other card, our category): 13 x 13 x 9 counters of 4 bits (<= 3 cards of a rank lie
                                       // outside a suit), eight per word
    uint32_t A0[104];                  // runouts WITHOUT a card of the suit holding >= 3 board cards, by (rank pair, our category):
                                       // 91 x 9 counters of 4 bits (<= 9 such card pairs per rank pair)
    uint32_t L1[2][512];               // non-zero entries of A1 / A0: key | multiplicity << 16 -- the lanes of those runout classes
    uint32_t L0[320];
    uint16_t pairs[kMaxQ5 + 8];        // colex pairs i | j << 8 (static)
    alignas(16) uint8_t T[kTRow5];     // [q][pp]: NF | keep(NF) << 4 -- one row of the global table of this board size
    uint8_t ourQ[kMaxQ5 + 8];          // our category on runout (i,j), colex index over the positions of D
    uint8_t clsP[kMaxP5 + 7];          // class of holding (i,j), colex index over the positions of U
    uint8_t clsR[96];                  // per rank pair: class from ranks alone
    uint8_t nf5[96];                   // per rank pair: NF(board + pair) | keep(NF) << 4
    uint8_t ppxy[96];                  // rank pair -> x | y << 4 (static)
    uint8_t code[56];                  // D = deck \ board: positions 0..m-1 the unseen cards (rank*4+suit ascending), m, m+1 our hole cards
    uint8_t Ls[4][20];                 // positions of D with suit s: unseen cards first, then our hole cards of that suit
    uint8_t Lns[4][52];                // positions of D with another suit, same order
    uint8_t nsu[4], nsd[4];            // cards of suit s among the unseen cards / among D
    uint8_t u[16];                     // unseen cards per rank
    uint8_t rstart[16];                // first position of each rank among the unseen cards
    uint32_t smask[4];                 // rank mask of the unseen cards per suit
};

__device__ __forceinline__ LitHand one_card5(uint32_t c) {
    LitHand h = {0, 0, 0, 0, 0};
    lit_add_card(h, c);
    return h;
}

// index of a sorted rank multiset r0 <= r1 <= ... in the combinatorial number system: sum of C(r_i + i, i + 1)
__device__ __forceinline__ uint32_t multiset_index(const uint32_t *r, int k) {
    uint32_t idx = 0;
    for (int i = 0; i < k; ++i) {
        uint32_t c = 1;                       // C(r_i + i, i + 1)
        for (int j = 0; j <= i; ++j) c = c * (r[i] + i - j) / (j + 1);
        idx += c;
    }
    return idx;
}

}  // namespace

// T tables of HOLDEM_MODE_REFERENCE, one row of kTRow5 bytes per board rank multiset (k = 3, 4, 5 board cards: 455 / 1820 / 6188
// rows): entry [q][pp] = category from the rank counts of board + rank pair q + rank pair pp (low nibble) and the value it takes
// when a flush that is no straight flush is present (high nibble).  Built on the device at holdem_init by the same arithmetic
// ranker the batched category kernel uses (pinned to the unmodified reference by tests/golden/literal_category.json).
__global__ void __launch_bounds__(256)
ref_t_table_kernel(uint8_t *__restrict__ table, int k, uint32_t n_boards) {
    __shared__ uint32_t ranks[5];
    __shared__ uint8_t ppxy[96];
    for (uint32_t b = blockIdx.x; b < n_boards; b += gridDim.x) {
        __syncthreads();
        if (threadIdx.x == 0) {               // unrank b: largest r with C(r + i, i + 1) <= rest, from the top position down
            uint32_t rest = b;
            for (int i = k - 1; i >= 0; --i) {
                uint32_t r = 0;
                for (;;) {
                    uint32_t c = 1;
                    for (int j = 0; j <= i; ++j) c = c * (r + 1 + i - j) / (j + 1);     // C(r + 1 + i, i + 1)
                    if (c > rest) break;
                    ++r;
                }
                uint32_t c = 1;
                for (int j = 0; j <= i; ++j) c = c * (r + i - j) / (j + 1);
                rest -= c;
                ranks[i] = r;
            }
        }
        if (threadIdx.x < 96) {
            uint32_t y = 0;
            while (tri(y + 1) <= threadIdx.x) ++y;
            ppxy[threadIdx.x] = threadIdx.x < (uint32_t)kPP5 ? (uint8_t)((threadIdx.x - tri(y)) | (y << 4)) : (uint8_t)0xFF;
        }
        __syncthreads();
        LitHand lb = {0, 0, 0, 0, 0};
        for (int i = 0; i < k; ++i) { lb.rc += 1ull << (4 * ranks[i]); lb.rmask |= 1u << ranks[i]; }
        uint8_t *row = table + (size_t)b * kTRow5;
        for (uint32_t e = threadIdx.x; e < (uint32_t)(kPP5 * kPP5); e += blockDim.x) {
            const uint32_t q = e / kPP5, p = e - q * kPP5;
            const uint32_t qxy = ppxy[q], pxy = ppxy[p];
            LitHand h = lb;
            const uint32_t r4[4] = {qxy & 15u, qxy >> 4, pxy & 15u, pxy >> 4};
#pragma unroll
            for (int t = 0; t < 4; ++t) { h.rc += 1ull << (4 * r4[t]); h.rmask |= 1u << r4[t]; }
            const uint32_t nf = lit_category(h);
            row[q * kTPitch5 + p] = (uint8_t)(nf | (((nf == 4u || nf >= 6u) ? nf : 5u) << 4));
        }
        for (uint32_t e = threadIdx.x; e < (uint32_t)kTRow5; e += blockDim.x)
            if (e >= (uint32_t)(kPP5 * kTPitch5) || e % kTPitch5 == (uint32_t)kPP5) row[e] = 0;     // padding bytes
    }
}

size_t ref_t_table_bytes(int k) { return (size_t)(k == 3 ? 455 : (k == 4 ? 1820 : 6188)) * kTRow5; }

cudaError_t launch_ref_t_table(uint8_t *table, int k, int sms, cudaStream_t st) {
    const uint32_t n_boards = k == 3 ? 455u : (k == 4 ? 1820u : 6188u);
    count_launch(); ref_t_table_kernel<<<sms * 4, 256, 0, st>>>(table, k, n_boards);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(kR5Threads, HOLDEM_R5_CTAS)
hp_ref5_kernel(const uint8_t *__restrict__ sit, int64_t n, uint32_t *__restrict__ out, DevTables T, uint32_t one,
               int *__restrict__ bad_flag, unsigned int *__restrict__ work_counter, PeerOut peers) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    Ref5Smem &S = *reinterpret_cast<Ref5Smem *>(smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // ---- static tables ---------------------------------------------------------------------------------------------------------
    for (uint32_t i = tid; i < (uint32_t)kMaxQ5; i += kR5Threads) S.pairs[i] = __ldg(T.pairs + i);
    {
        uint32_t word = 0;
        for (uint32_t b = 0; b < 32; ++b) word |= (uint32_t)cyclic_run5(tid * 32 + b) << b;
        S.cyc[tid] = word;
    }
    if (tid < 96) {
        uint32_t y = 0;
        while (tri(y + 1) <= tid) ++y;
        S.ppxy[tid] = tid < kPP5 ? (uint8_t)((tid - tri(y)) | (y << 4)) : (uint8_t)0xFF;
    }
    const uint32_t t_s = (uint32_t)__cvta_generic_to_shared(S.T);
    const uint32_t cyc_s = (uint32_t)__cvta_generic_to_shared(S.cyc);
    const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(&S.bar);
    if (tid == 0) {
        mbar_init(bar_s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t parity = 0;

    for (;;) {
        __syncthreads();                                                                            // previous situation consumed
        if (tid == 0) S.claim = atomicAdd(work_counter, 1u);
        __syncthreads();
        const int64_t idx = S.claim;
        if (idx >= n) break;
        // ---- parse -----------------------------------------------------------------------------------------------------------------
        const uint2 w = __ldg(reinterpret_cast<const uint2 *>(sit) + idx);
        uint32_t cb[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) { cb[k] = (w.x >> (8 * k)) & 0xFFu; cb[4 + k] = (w.y >> (8 * k)) & 0xFFu; }
        int nb = 0;
        bool ok = cb[0] < 52u && cb[1] < 52u, ended = false;
        uint64_t known = card_bit(cb[0] < 52u ? cb[0] : 63u) | card_bit(cb[1] < 52u ? cb[1] : 63u), bset = 0;
        LitHand lb = {0, 0, 0, 0, 0};
        uint32_t bsuit = 0;                      // board cards per suit, 4 bits each
#pragma unroll
        for (int k = 2; k < 7; ++k) {
            const uint32_t c = cb[k];
            if (c == 0xFFu) ended = true;
            else {
                ok = ok && !ended && c < 52u;
                const uint32_t cc = c < 52u ? c : 0u;
                nb += 1;
                bset |= card_bit(cc);
                lit_add_card(lb, cc);
                bsuit += 1u << (4 * card_suit(cc));
            }
        }
        known |= bset;
        ok = ok && nb >= 3 && __popcll(known) == nb + 2;
        if (!ok) {
            if (tid == 0 && bad_flag != nullptr) atomicOr(bad_flag, 4);   // bit 2: a malformed row was marked
            if (tid < 16) {
                out[idx * 16 + tid] = 0xFFFFFFFFu;
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (k < peers.n) peers.p[k][idx * 16 + tid] = 0xFFFFFFFFu;
            }
            continue;
        }
        const uint32_t m = 50u - (uint32_t)nb;          // unseen cards
        const uint32_t md = m + 2u;                     // cards of D = deck \ board
        const uint32_t n_p = m * (m - 1) / 2, n_q = md * (md - 1) / 2;
        const LitHand lours = lit_merge(lb, lit_merge(one_card5(cb[0]), one_card5(cb[1])));
        const uint32_t ours_now = lit_category(lours);                                           // :190
        const uint32_t hr0 = card_rank(cb[0]), hr1 = card_rank(cb[1]);
        const uint32_t hole_pp = tri(max(hr0, hr1)) + min(hr0, hr1);
        const uint32_t hsuit = bsuit + (1u << (4 * card_suit(cb[0]))) + (1u << (4 * card_suit(cb[1])));   // hole + board per suit
        const uint32_t h5 = (hsuit + 0x3333u) & 0x8888u;                                         // we already hold five of a suit
        if (tid < 12) S.cnt[tid] = 0;
        if (tid < 4) S.bmask[tid] = (uint32_t)(bset >> (13 * tid)) & 0x1FFFu;
        for (uint32_t i = tid; i < (uint32_t)(kPP5 * 9); i += kR5Threads) S.M[i] = 0;
        // Suits with >= 2 board cards (at most two) own a slot of A1 / L1, the suit with >= 3 (at most one) owns A0 / L0.
        uint32_t slotmap = 0, n_slots = 0, a0suit = 4;
#pragma unroll
        for (uint32_t su = 0; su < 4; ++su) {
            const uint32_t bsq = (bsuit >> (4 * su)) & 15u;
            slotmap |= (bsq >= 2 ? n_slots++ : 15u) << (4 * su);
            if (bsq >= 3) a0suit = su;
        }
        if (n_slots) for (uint32_t i = tid; i < n_slots * 192u; i += kR5Threads) (&S.A1[0][0])[i] = 0;
        if (a0suit < 4) for (uint32_t i = tid; i < 104u; i += kR5Threads) S.A0[i] = 0;
        if (tid < 3) (&S.nL1[0])[tid] = 0;          // nL1[0], nL1[1], nL0
        if (tid == 3) S.n_entries = 0;
        if (warp == 0) {
            // ---- D in (rank, suit) order, our hole cards last; per-suit position lists ----------------------------------------------
            const uint32_t p0 = path_of_code4(lane), p1 = path_of_code4(lane + 32);
            const bool in0 = !((known >> p0) & 1ull);
            const bool in1 = (lane + 32 < 52) && !((known >> p1) & 1ull);
            const uint32_t m0 = __ballot_sync(0xFFFFFFFFu, in0), m1 = __ballot_sync(0xFFFFFFFFu, in1);
            const uint32_t ltm = (1u << lane) - 1u;
            if (in0) S.code[__popc(m0 & ltm)] = (uint8_t)lane;
            if (in1) S.code[__popc(m0) + __popc(m1 & ltm)] = (uint8_t)(lane + 32);
            if (lane < 2) {
                const uint32_t h = cb[lane];
                S.code[m + lane] = (uint8_t)(card_rank(h) * 4 + card_suit(h));
            }
            if (lane < 13) {
                const uint32_t c4 = 4 * lane;
                S.u[lane] = (uint8_t)(c4 < 32 ? __popc((m0 >> c4) & 15u) : __popc((m1 >> (c4 - 32)) & 15u));
                S.rstart[lane] = (uint8_t)(c4 < 32 ? __popc(m0 & ((1u << c4) - 1u)) : __popc(m0) + __popc(m1 & ((1u << (c4 - 32)) - 1u)));
            }
            __syncwarp();
            const uint32_t cd0 = S.code[lane], cd1 = lane + 32 < md ? S.code[lane + 32] : 0u;
            const bool v1 = lane + 32 < md;
#pragma unroll
            for (uint32_t s = 0; s < 4; ++s) {
                const bool a0 = (cd0 & 3u) == s, a1 = v1 && (cd1 & 3u) == s;
                const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, a0), b1 = __ballot_sync(0xFFFFFFFFu, a1);
                const uint32_t vb1 = __ballot_sync(0xFFFFFFFFu, v1);
                const uint32_t n0 = __popc(b0);
                if (a0) S.Ls[s][__popc(b0 & ltm)] = (uint8_t)lane;
                else S.Lns[s][lane - __popc(b0 & ltm)] = (uint8_t)lane;
                if (v1) {
                    if (a1) S.Ls[s][n0 + __popc(b1 & ltm)] = (uint8_t)(lane + 32);
                    else S.Lns[s][(32 - n0) + __popc(vb1 & ~b1 & ltm)] = (uint8_t)(lane + 32);
                }
                const bool un1 = lane + 32 < m;                       // unseen (not one of our hole cards)
                const uint32_t rm = __reduce_or_sync(0xFFFFFFFFu, (a0 ? 1u << (cd0 >> 2) : 0u) | ((a1 && un1) ? 1u << (cd1 >> 2) : 0u));
                if (lane == 0) {
                    S.smask[s] = rm;
                    const uint32_t nd = n0 + __popc(b1);
                    const uint32_t hole_s = (uint32_t)((S.code[m] & 3u) == s) + (uint32_t)((S.code[m + 1] & 3u) == s);
                    S.nsd[s] = (uint8_t)nd;
                    S.nsu[s] = (uint8_t)(nd - hole_s);     // positions ascend, the hole cards sit last: unseen cards come first
                }
            }
        }
        if (tid == 32) {
            // ---- T = the board's row of the global table (8,384 bytes), fetched by the TMA engine under the set-up phases -------------
            uint32_t br[5];
            int k = 0;
#pragma unroll
            for (int j = 2; j < 7; ++j)
                if (j - 2 < nb) br[k++] = card_rank(cb[j]);
            for (int i = 1; i < nb; ++i)                               // insertion sort, <= 5 ranks
                for (int j = i; j > 0 && br[j - 1] > br[j]; --j) { const uint32_t t = br[j]; br[j] = br[j - 1]; br[j - 1] = t; }
            const uint32_t bi = multiset_index(br, nb);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier reads of T precede the refill
            mbar_expect_tx(bar_s, kTRow5);
            bulk_g2s(t_s, T.ref_t[nb - 3] + (size_t)bi * kTRow5, kTRow5, bar_s);
        }
        if (tid >= 64 && tid < 160) {
            // ---- per rank pair: the category board + pair takes from its ranks alone, and its class against ours -----------------------------
            const uint32_t pp = tid - 64;
            uint32_t clsr = 3, e8 = 0;
            if (pp < (uint32_t)kPP5) {
                const uint32_t xy = S.ppxy[pp], x = xy & 15u, y = xy >> 4;
                LitHand h = lb;
                h.sc = 0; h.cards = 0; h.dup = 0;
                h.rc += (1ull << (4 * x)) + (1ull << (4 * y));
                h.rmask |= (1u << x) | (1u << y);
                const uint32_t nf = lit_category(h);
                clsr = ours_now > nf ? 0u : (ours_now == nf ? 1u : 2u);
                e8 = nf | (((nf == 4u || nf >= 6u) ? nf : 5u) << 4);
            }
            S.clsR[pp] = (uint8_t)clsr;
            S.nf5[pp] = (uint8_t)e8;
        }
        __syncthreads();                                                                            // (A) lists, u[], clsR, nf5 ready
        if (tid < 96) {                 // holdings per rank pair, packed by class
            uint32_t wt = 0;
            if (tid < (uint32_t)kPP5) {
                const uint32_t xy = S.ppxy[tid], x = xy & 15u, y = xy >> 4;
                const uint32_t ux = S.u[x], uy = S.u[y];
                const uint32_t cntp = x != y ? ux * uy : (ux ? ux * (ux - 1) / 2 : 0u);
                wt = cntp << (10 * S.clsR[tid]);
            }
            S.Wt32[tid] = wt;
        }
        // ---- class of every holding (:194-203): the class of its rank pair unless board + holding hold five of a suit -------------------
        {
            uint32_t c0 = 0, c1 = 0, c2 = 0;
            for (uint32_t p = tid; p < n_p; p += kR5Threads) {
                const uint32_t pr = S.pairs[p];
                const uint32_t ca = S.code[pr & 0xFFu], cb2 = S.code[pr >> 8];
                const uint32_t ra = ca >> 2, rb = cb2 >> 2, sa = ca & 3u, sb = cb2 & 3u;
                const uint32_t pp = tri(max(ra, rb)) + min(ra, rb);
                uint32_t cls = S.clsR[pp];
                // suit counts of board + holding (distinct cards): a flush needs 5 - (board cards of the suit) cards of it in the holding
                const uint32_t na = ((bsuit >> (4 * sa)) & 15u) + 1u + (uint32_t)(sb == sa), nbs = ((bsuit >> (4 * sb)) & 15u) + 1u;
                const uint32_t f5 = (bsuit + 0x3333u) & 0x8888u;                    // a suit with five board cards
                uint32_t fs = 4;
                if (na >= 5) fs = sa;
                else if (nbs >= 5) fs = sb;
                else if (f5) fs = (uint32_t)(__ffs(f5) - 1) >> 2;
                if (fs < 4) {
                    const uint32_t mask = S.bmask[fs] | (sa == fs ? 1u << ra : 0u) | (sb == fs ? 1u << rb : 0u);
                    const uint32_t e8 = S.nf5[pp];
                    const uint32_t opp = ((S.cyc[mask >> 5] >> (mask & 31u)) & 1u) ? 8u : (e8 >> 4);
                    cls = ours_now > opp ? 0u : (ours_now == opp ? 1u : 2u);
                }
                S.clsP[p] = (uint8_t)cls;
                c0 += cls == 0; c1 += cls == 1; c2 += cls == 2;
            }
            c0 = __reduce_add_sync(0xFFFFFFFFu, c0);
            c1 = __reduce_add_sync(0xFFFFFFFFu, c1);
            c2 = __reduce_add_sync(0xFFFFFFFFu, c2);
            if (lane == 0) { atomicAdd(&S.cnt[6], c0); atomicAdd(&S.cnt[7], c1); atomicAdd(&S.cnt[8], c2); }
        }
        mbar_wait(bar_s, parity);                                      // T has landed
        parity ^= 1u;
        // ---- our category on every runout (:207-210; a runout may repeat our hole cards) ---------------------------------------------
        for (uint32_t q = tid; q < n_q; q += kR5Threads) {
            const uint32_t pr = S.pairs[q];
            const uint32_t pi = pr & 0xFFu, pj = pr >> 8;
            const uint32_t ci = S.code[pi], cj = S.code[pj];
            const uint32_t ri = ci >> 2, rj = cj >> 2, si = ci & 3u, sj = cj & 3u;
            const uint32_t qp = tri(max(ri, rj)) + min(ri, rj);
            // hole + board + runout: the category of its rank counts is T[rank pair of our hole cards][rank pair of the runout];
            // a flush needs five of a suit counted WITH repeats (a runout card that is one of our hole cards counts twice, :56)
            const uint32_t e8 = S.T[hole_pp * kTPitch5 + qp];
            uint32_t o = e8 & 15u;
            const uint32_t ni = ((hsuit >> (4 * si)) & 15u) + 1u + (uint32_t)(sj == si), nj = ((hsuit >> (4 * sj)) & 15u) + 1u;
            uint32_t fs = 4;
            if (ni >= 5) fs = si;
            else if (nj >= 5) fs = sj;
            else if (h5) fs = (uint32_t)(__ffs(h5) - 1) >> 2;
            if (fs < 4) {
                const uint32_t mask = ((uint32_t)(known >> (13 * fs)) & 0x1FFFu) | (si == fs ? 1u << ri : 0u) | (sj == fs ? 1u << rj : 0u);
                const bool repeated = (si == fs && pi >= m) || (sj == fs && pj >= m);     // one of our hole cards again: no straight flush
                o = (!repeated && ((S.cyc[mask >> 5] >> (mask & 31u)) & 1u)) ? 8u : (e8 >> 4);
            }
            S.ourQ[q] = (uint8_t)o;
            atomicAdd(&S.M[qp * 9 + o], 1u);
            // the runout as a lane of the flush part: with a card outside a suit it matters through that card's RANK only
            if (si != sj) {                                  // exactly one card in suit si and one in suit sj
                const uint32_t sl_i = (slotmap >> (4 * si)) & 15u, sl_j = (slotmap >> (4 * sj)) & 15u;
                // (key = our category x 169 + rank of the card in s x 13 + rank of the other card: category-major, see the lane lists)
                if (sl_i < 2) { const uint32_t k = o * 169u + ri * 13u + rj; atomicAdd(&S.A1[sl_i][k >> 3], 1u << (4 * (k & 7u))); }
                if (sl_j < 2) { const uint32_t k = o * 169u + rj * 13u + ri; atomicAdd(&S.A1[sl_j][k >> 3], 1u << (4 * (k & 7u))); }
            }
            if (a0suit < 4 && si != a0suit && sj != a0suit) {
                const uint32_t k = o * (uint32_t)kPP5 + qp;
                atomicAdd(&S.A0[k >> 3], 1u << (4 * (k & 7u)));
            }
        }
        if (warp == 0) {
            // ---- flush task groups and descriptor sections: lane = 3 * suit + (2 - qtype) ------------------------------------------------
            const uint32_t ltm = (1u << lane) - 1u;
            const uint32_t s = min(lane / 3u, 3u), qtype = 2u - (lane - 3u * (lane / 3u));
            const uint32_t bs = (bsuit >> (4 * s)) & 15u;
            const uint32_t nu = S.nsu[s], nnu = m - nu, nd = S.nsd[s], nnd = md - nd;
            // holdings of the suit by their cards in s: [two, card by card | one: (card a in s, rank of the other card) |
            // none: rank pairs] -- cards outside s matter through their rank only, so those holdings are merged into weighted steps
            const uint32_t len2 = choose2u(nu), len1 = len2 + 13u * nu, len0 = len1 + (uint32_t)kPP5;
            (void)nnu;
            // longest prefix any runout class needs: cards of s the holding must bring when the runout brings two
            const int need2 = 5 - (int)bs - 2;
            const uint32_t len = bs == 0 ? 0u : (need2 >= 3 ? 0u : (need2 == 2 ? len2 : (need2 == 1 ? len1 : len0)));
            uint32_t pdo = 0;
#pragma unroll
            for (uint32_t q = 0; q < 3; ++q) {
                const uint32_t lq = __shfl_sync(0xFFFFFFFFu, len, 3 * q);
                if (q < s) pdo += lq;
            }
            const int need = 5 - (int)bs - (int)qtype;                 // cards of s the holding must bring for this runout class
            const uint32_t np = (bs == 0 || need >= 3) ? 0u : (need == 2 ? len2 : (need == 1 ? len1 : len0));
            (void)nd; (void)nnd; (void)ltm;
            if (lane < 12) { S.h_np[lane] = np; S.h_pdo[lane] = pdo; }
            if (lane < 12 && qtype == 2) { S.pd_base[s] = pdo; S.pd_len[s] = len; }
        }
        __syncthreads();                                                                            // (B)
        // ---- compact the (rank pair, our category) entries that occur ---------------------------------------------------------------
        for (uint32_t e = tid; e < (uint32_t)(kPP5 * 9 + kR5Threads - 1) / kR5Threads * kR5Threads; e += kR5Threads) {
            const bool nz = e < (uint32_t)(kPP5 * 9) && S.M[e] != 0;
            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, nz);
            uint32_t base = 0;
            if (lane == 0 && bal) base = atomicAdd(&S.n_entries, __popc(bal));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (nz) S.elist[base + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)e;
        }
        // ---- the merged runout lanes: non-zero counters of A1 (per slot) and A0, as key | multiplicity << 16 ------------------------------
        {
            // The keys are category-major and the lists are filled in two passes, categories 0..3 (three of a kind or less) first:
            // a warp task whose 32 lanes all hold such a category loses to every flush step's add part (the opponent's category is
            // 4..8 there) and walks the undo half alone.
            auto compact16 = [&](const uint32_t *arr, uint32_t k0, uint32_t k1, uint32_t *list, uint32_t *counter) {
                for (uint32_t e = k0 + tid; e < k0 + (k1 - k0 + kR5Threads - 1) / kR5Threads * kR5Threads; e += kR5Threads) {
                    const uint32_t c = e < k1 ? (arr[e >> 3] >> (4 * (e & 7u))) & 15u : 0u;
                    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, c != 0);
                    uint32_t base = 0;
                    if (lane == 0 && bal) base = atomicAdd(counter, __popc(bal));
                    base = __shfl_sync(0xFFFFFFFFu, base, 0);
                    if (c) list[base + __popc(bal & ((1u << lane) - 1u))] = e | (c << 16);
                }
            };
            for (uint32_t sl = 0; sl < n_slots; ++sl) compact16(S.A1[sl], 0u, 4u * 169u, S.L1[sl], &S.nL1[sl]);
            if (a0suit < 4) compact16(S.A0, 0u, 4u * (uint32_t)kPP5, S.L0, &S.nL0);
            __syncthreads();
            for (uint32_t sl = 0; sl < n_slots; ++sl) compact16(S.A1[sl], 4u * 169u, 9u * 169u, S.L1[sl], &S.nL1[sl]);
            if (a0suit < 4) compact16(S.A0, 4u * (uint32_t)kPP5, 9u * (uint32_t)kPP5, S.L0, &S.nL0);
        }
        // ---- holding descriptors of the flush part --------------------------------------------------------------------------------------
        {
            const uint32_t total = S.pd_base[3] + S.pd_len[3];
            for (uint32_t e = tid; e < total; e += kR5Threads) {
                uint32_t s = 0;
#pragma unroll
                for (uint32_t q = 1; q < 4; ++q)
                    if (S.pd_len[q] && e >= S.pd_base[q]) s = q;
                const uint32_t t = e - S.pd_base[s];
                const uint32_t nu = S.nsu[s], nboth = choose2u(nu), sm = S.smask[s];
                // first / second unseen card of rank r outside s (positions; the unseen cards are sorted by rank, then suit)
                auto rep = [&](uint32_t r, uint32_t skip) -> uint32_t {
                    uint32_t pos = S.rstart[r];
                    for (uint32_t k = 0; k < 4u; ++k) {
                        const uint32_t ps = pos + k;
                        if ((S.code[ps] & 3u) != s) { if (skip == 0) return ps; --skip; }
                    }
                    return pos;
                };
                uint32_t bits, pp, wgt, pa, pb;                    // pa, pb: a representative holding (positions in D, both < m)
                if (t < nboth) {                                   // both cards in s: card by card
                    const uint32_t pr = S.pairs[t];
                    pa = S.Ls[s][pr & 0xFFu]; pb = S.Ls[s][pr >> 8];
                    const uint32_t ra = S.code[pa] >> 2, rb = S.code[pb] >> 2;
                    bits = (1u << ra) | (1u << rb);
                    pp = tri(max(ra, rb)) + min(ra, rb);
                    wgt = 1;
                } else if (t < nboth + 13u * nu) {                 // card a in s, any unseen card of rank y outside s
                    const uint32_t t2 = t - nboth, ai = t2 / 13u, y = t2 - 13u * ai;
                    pa = S.Ls[s][ai];
                    const uint32_t ra = S.code[pa] >> 2;
                    bits = 1u << ra;
                    pp = tri(max(ra, y)) + min(ra, y);
                    wgt = S.u[y] - ((sm >> y) & 1u);
                    pb = wgt ? rep(y, 0) : pa;
                } else {                                           // a rank pair, neither card in s
                    pp = t - nboth - 13u * nu;
                    const uint32_t xy = S.ppxy[pp], x = xy & 15u, y = xy >> 4;
                    const uint32_t wx = S.u[x] - ((sm >> x) & 1u), wy = S.u[y] - ((sm >> y) & 1u);
                    bits = 0;
                    wgt = x != y ? wx * wy : (wx ? wx * (wx - 1) / 2 : 0u);
                    pa = wgt ? rep(x, 0) : 0u;
                    pb = wgt ? (x != y ? rep(y, 0) : rep(x, 1)) : 1u;
                }
                const uint32_t lo = min(pa, pb), hi = max(pa, pb);
                const uint32_t cls = wgt && lo != hi ? (uint32_t)S.clsP[hi * (hi - 1) / 2 + lo] : 0u;
                S.PD[e] = make_uint4(bits, wgt << (10 * cls), wgt << (10 * S.clsR[pp]), pp);
            }
        }
        __syncthreads();                                                                            // (C)

        if (warp == 0) {
            // ---- live flush task groups: lane = 3 * suit + (2 - qtype); runouts with two cards in s are lanes one by one, the others
            //      are the merged lanes of L1 / L0 ------------------------------------------------------------------------------------------
            const uint32_t ltm = (1u << lane) - 1u;
            const uint32_t s = min(lane / 3u, 3u), qtype = 2u - (lane - 3u * (lane / 3u));
            const uint32_t slot = (slotmap >> (4 * s)) & 15u;
            const uint32_t np = lane < 12 ? S.h_np[lane] : 0u, pdo = lane < 12 ? S.h_pdo[lane] : 0u;
            const uint32_t nq = qtype == 2 ? choose2u(S.nsd[s]) : (qtype == 1 ? (slot < 2 ? S.nL1[slot] : 0u) : (s == a0suit ? S.nL0 : 0u));
            const bool live = lane < 12 && nq != 0 && np != 0;
            const uint32_t nseg = (np + kSeg5 - 1) / kSeg5;
            const uint32_t ntask = live ? ((nq + 31) / 32) * nseg : 0u;
            const uint32_t livem = __ballot_sync(0xFFFFFFFFu, live);
            const uint32_t gi = __popc(livem & ltm);
            uint32_t first = ntask;
#pragma unroll
            for (int d = 1; d < 16; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, first, d);
                if ((int)lane >= d) first += v;
            }
            if (live) {
                S.g_first[gi] = first - ntask;
                S.g_nq[gi] = nq; S.g_np[gi] = np; S.g_nseg[gi] = nseg; S.g_pdbase[gi] = pdo;
                S.g_sq[gi] = s | (qtype << 2) | (slot << 4);
            }
            if (lane == 11) {
                S.g_first[__popc(livem)] = first;
                S.n_groups = __popc(livem);
                S.next_task = 0;
            }
        }
        uint32_t accA[3] = {0, 0, 0}, accB[3] = {0, 0, 0};     // per class: our > opp (ahead) / our < opp (behind), net, mod 2^32
        // ---- generic part: eight lanes share one (rank pair, our category) entry, each takes every eighth rank pair (<= 12 of them:
        //      a packed 10-bit field holds at most 12 x 16 holdings); the fields are spread to 16 bits before the lanes are summed ----------
        {
            const uint32_t n_entries = S.n_entries;
            const uint32_t sub = lane & 7u;
            for (uint32_t k0 = tid >> 3; k0 < ((n_entries + 3u) & ~3u); k0 += kR5Threads / 8) {   // warp-uniform trip count
                const bool act = k0 < n_entries;
                const uint32_t e = act ? (uint32_t)S.elist[k0] : 0u, q = e / 9u, o = e - 9u * q;
                const uint8_t *row = S.T + q * kTPitch5;
                uint32_t gt = 0, lt = 0;
                for (uint32_t pp = sub; pp < (uint32_t)kPP5; pp += 8) {
                    const uint32_t nf = (uint32_t)row[pp] & 15u, w = S.Wt32[pp];
                    if (o > nf) gt += w;
                    if (o < nf) lt += w;
                }
                uint32_t g01 = (gt & 0x3FFu) | (((gt >> 10) & 0x3FFu) << 16), g2 = gt >> 20;
                uint32_t l01 = (lt & 0x3FFu) | (((lt >> 10) & 0x3FFu) << 16), l2 = lt >> 20;
#pragma unroll
                for (int d = 1; d < 8; d <<= 1) {
                    g01 += __shfl_xor_sync(0xFFFFFFFFu, g01, d); g2 += __shfl_xor_sync(0xFFFFFFFFu, g2, d);
                    l01 += __shfl_xor_sync(0xFFFFFFFFu, l01, d); l2 += __shfl_xor_sync(0xFFFFFFFFu, l2, d);
                }
                if (act && sub == 0) {
                    const uint32_t mult = S.M[e];
                    accA[0] += (g01 & 0xFFFFu) * mult; accA[1] += (g01 >> 16) * mult; accA[2] += g2 * mult;
                    accB[0] += (l01 & 0xFFFFu) * mult; accB[1] += (l01 >> 16) * mult; accB[2] += l2 * mult;
                }
            }
        }
        __syncthreads();                                                                            // (C2) task groups ready
        // ---- flush part: dynamic task queue ----------------------------------------------------------------------------------------------------
        {
            const uint32_t n_groups = S.n_groups;
            const uint32_t n_tasks = S.g_first[n_groups];
            for (;;) {
                uint32_t task = 0;
                if (lane == 0) task = atomicAdd(&S.next_task, 1u);
                task = __shfl_sync(0xFFFFFFFFu, task, 0);
                if (task >= n_tasks) break;
                uint32_t g = 0;
                while (g + 1 < n_groups && task >= S.g_first[g + 1]) ++g;
                const uint32_t local = task - S.g_first[g];
                const uint32_t nseg = S.g_nseg[g], np = S.g_np[g], nq = S.g_nq[g];
                const uint32_t chunk = local / nseg, seg = local - chunk * nseg;
                const uint32_t s = S.g_sq[g] & 3u, qtype = (S.g_sq[g] >> 2) & 3u, slot = S.g_sq[g] >> 4;
                const uint32_t p0 = seg * kSeg5, p1 = min(np, p0 + kSeg5);
                const uint32_t qi = chunk * 32 + lane;
                uint32_t our = 0, qbits = 0, trow = 0, mult = 0;
                const bool valid = qi < nq;
                if (valid) {
                    if (qtype == 2) {                              // both cards in s: the runout itself
                        const uint32_t pr = S.pairs[qi];
                        const uint32_t x = S.Ls[s][pr & 0xFFu], y = S.Ls[s][pr >> 8];     // positions in D, x < y
                        const uint32_t rx = S.code[x] >> 2, ry = S.code[y] >> 2;
                        our = S.ourQ[y * (y - 1) / 2 + x];
                        qbits = (1u << rx) | (1u << ry);
                        trow = (tri(max(rx, ry)) + min(rx, ry)) * kTPitch5;
                        mult = 1;
                    } else if (qtype == 1) {                       // (card of s by its rank, rank of the other card, our category)
                        const uint32_t en = S.L1[slot][qi], key = en & 0xFFFFu;
                        our = key / 169u;
                        const uint32_t rem = key - 169u * our, rx = rem / 13u, ry = rem - 13u * rx;
                        qbits = 1u << rx;
                        trow = (tri(max(rx, ry)) + min(rx, ry)) * kTPitch5;
                        mult = en >> 16;
                    } else {                                       // (rank pair, our category), neither card in s
                        const uint32_t en = S.L0[qi], key = en & 0xFFFFu;
                        our = key / (uint32_t)kPP5;
                        const uint32_t qp = key - (uint32_t)kPP5 * our;
                        trow = qp * kTPitch5;
                        mult = en >> 16;
                    }
                }
                const uint32_t base_mask = S.bmask[s] | qbits;
                const uint32_t tq_s = t_s + trow;
                const uint4 *pd = S.PD + S.g_pdbase[g] + p0;
                uint32_t aG = 0, aL = 0, sG = 0, sL = 0;
                if (__all_sync(0xFFFFFFFFu, our <= 3u)) {
                    // every lane holds three of a kind or less: the add part of every step is a loss for us (4 <= keep(NF), 8), so
                    // the steps carry only their undo half -- no cyclic-run test, no category select, half the compares
#pragma unroll 4
                    for (uint32_t t = p0; t < p1; ++t, ++pd) {
                        const uint4 d = *pd;
                        uint32_t e8;
                        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(e8) : "r"(tq_s + d.w));
                        const uint32_t catn = e8 & 15u;
                        aL += d.y;
                        asm("{\n\t.reg .pred p;\n\t"
                            "setp.gt.u32 p, %2, %3;\n\t@p mad.lo.u32 %0, %4, %5, %0;\n\t"
                            "setp.lt.u32 p, %2, %3;\n\t@p mad.lo.u32 %1, %4, %5, %1;\n\t}"
                            : "+r"(sG), "+r"(sL)
                            : "r"(our), "r"(catn), "r"(d.z), "r"(one));
                    }
                } else
#pragma unroll 4
                for (uint32_t t = p0; t < p1; ++t, ++pd) {
                    const uint4 d = *pd;
                    const uint32_t pbits = d.x;
                    uint32_t e8;
                    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(e8) : "r"(tq_s + d.w));
                    const uint32_t mask = base_mask | pbits;
                    uint32_t cw;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(cw) : "r"(cyc_s + ((mask >> 5) << 2)));
                    const bool sf = ((cw >> (mask & 31u)) & 1u) != 0 && (qbits & pbits) == 0;   // no shared card of s, cyclic run
                    const uint32_t catf = sf ? 8u : (e8 >> 4), catn = e8 & 15u;
                    // if (our > catf) aG += wa; if (our < catf) aL += wa; if (our > catn) sG += ws; if (our < catn) sL += ws
                    asm("{\n\t.reg .pred p;\n\t"
                        "setp.gt.u32 p, %4, %5;\n\t@p mad.lo.u32 %0, %7, %9, %0;\n\t"
                        "setp.lt.u32 p, %4, %5;\n\t@p mad.lo.u32 %1, %7, %9, %1;\n\t"
                        "setp.gt.u32 p, %4, %6;\n\t@p mad.lo.u32 %2, %8, %9, %2;\n\t"
                        "setp.lt.u32 p, %4, %6;\n\t@p mad.lo.u32 %3, %8, %9, %3;\n\t}"
                        : "+r"(aG), "+r"(aL), "+r"(sG), "+r"(sL)
                        : "r"(our), "r"(catf), "r"(catn), "r"(d.y), "r"(d.z), "r"(one));
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) {       // mult = 0 on lanes past the end of the lane list
                    accA[c] += (((aG >> (10 * c)) & 0x3FFu) - ((sG >> (10 * c)) & 0x3FFu)) * mult;
                    accB[c] += (((aL >> (10 * c)) & 0x3FFu) - ((sL >> (10 * c)) & 0x3FFu)) * mult;
                }
            }
        }
        // ---- reduce over the warp, then into the CTA's counters ------------------------------------------------------------------------------
#pragma unroll
        for (uint32_t c = 0; c < 3; ++c) {
            const uint32_t a = __reduce_add_sync(0xFFFFFFFFu, accA[c]);
            const uint32_t b = __reduce_add_sync(0xFFFFFFFFu, accB[c]);
            if (lane == 0) {
                if (a) atomicAdd(&S.cnt[3 + c], a);
                if (b) atomicAdd(&S.cnt[c], b);
            }
        }
        __syncthreads();                                                                            // (D)
        if (tid < 16) {
            uint32_t v;
            if (tid < 3) v = S.cnt[6 + tid];
            else if (tid < 12) {
                const uint32_t cls = (tid - 3) / 3, o = (tid - 3) % 3;
                const uint32_t behind = S.cnt[cls], ahead = S.cnt[3 + cls], all = n_q * S.cnt[6 + cls];
                v = o == 0 ? ahead : (o == 2 ? behind : all - ahead - behind);   // HP[class][ahead|tied|behind] (:213-218)
            } else if (tid < 15) v = S.cnt[6 + tid - 12];
            else v = (uint32_t)nb;
            out[idx * 16 + tid] = v;
#pragma unroll
            for (int k = 0; k < 8; ++k)          // other GPUs' result buffers (NVLink stores)
                if (k < peers.n) peers.p[k][idx * 16 + tid] = v;
        }
    }
}

cudaError_t launch_hp_ref5(const DevTables &T, const uint8_t *sit, int64_t n, uint32_t *out, int sms, int *flag_slot,
                           const PeerOut &peers, cudaStream_t st) {
    // flag_slot points at the call's 16-byte slot: [0] flag word (bit 2 = malformed row marked), [1] this kernel's work counter
    if (n == 0) return cudaSuccess;
    if (n >= ((int64_t)1 << 31)) return cudaErrorInvalidValue;
    const size_t smem = sizeof(Ref5Smem);
    cudaError_t e = cudaFuncSetAttribute(hp_ref5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    static int bps = 0;
    if (bps == 0) {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, hp_ref5_kernel, kR5Threads, smem);
        if (e != cudaSuccess) return e;
        if (bps < 1) bps = 1;
    }
    int64_t blocks = (int64_t)sms * bps;
    if (blocks > n) blocks = n;
    e = cudaMemsetAsync(flag_slot, 0, 4 * sizeof(int), st);
    if (e != cudaSuccess) return e;
    count_launch(); hp_ref5_kernel<<<(int)blocks, kR5Threads, smem, st>>>(sit, n, out, T, 1u, flag_slot,
                                                                          reinterpret_cast<unsigned int *>(flag_slot) + 1, peers);
    return cudaGetLastError();
}

// =====================================================================================================================================
// Hand strength in HOLDEM_MODE_REFERENCE from rank pairs (HandCalculations.py:157-176 with the literal ranker): one warp per
// situation, no shared memory, no tables.
//   ahead / tied / behind over the C(m,2) opponent holdings.  A holding whose five / six / seven cards with the board hold no
//   five of a suit is ranked by its rank pair alone:  sum over the 91 rank pairs of n(pp) x cmp(ours, NF(board + pp)), three
//   rank pairs per lane, NF from the arithmetic ranker on rank counts.  Holdings that do make a flush with the board (the suit
//   with >= 3 board cards: both cards in it; with >= 4: one card in it, merged by the rank of the other; with 5: every holding,
//   merged by rank pair) move from the class of NF to the class of  8 if the suit's ranks hold a cyclic five-run, else NF when it
//   is 7, 6 or 4, else 5.
__global__ void __launch_bounds__(256)
hs_ref5_kernel(const uint8_t *__restrict__ sit, int64_t n, uint32_t *__restrict__ out) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n_warps = (int64_t)gridDim.x * 8;
    // the lane's three rank pairs x <= y (of 91) and its three pairs of distinct ranks x < y (of 78), packed x | y << 4
    uint32_t ppk[3], sqk[3];
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        const uint32_t pp = lane + 32u * p;
        uint32_t y = 0;
        while (tri(y + 1) <= pp) ++y;
        ppk[p] = pp < (uint32_t)kPP5 ? ((pp - tri(y)) | (y << 4)) : 0xFFu;
        y = 1;
        while (y * (y + 1) / 2 <= pp) ++y;
        sqk[p] = pp < 78u ? ((pp - y * (y - 1) / 2) | (y << 4)) : 0xFFu;
    }
    for (int64_t idx = (int64_t)blockIdx.x * 8 + warp; idx < n; idx += n_warps) {
        const uint2 w = __ldg(reinterpret_cast<const uint2 *>(sit) + idx);
        uint32_t cb[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) { cb[k] = (w.x >> (8 * k)) & 0xFFu; cb[4 + k] = (w.y >> (8 * k)) & 0xFFu; }
        const uint32_t nb = cb[5] == 0xFFu ? 3u : (cb[6] == 0xFFu ? 4u : 5u);
        bool ok = true;
        uint64_t known = 0, bset = 0;
        uint32_t bsuit = 0;
        LitHand lb = {0, 0, 0, 0, 0}, lh = {0, 0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            const uint32_t c = cb[k];
            const bool used = (uint32_t)k < nb + 2;
            ok = ok && (used ? c < 52u : c == 0xFFu);
            const uint32_t cc = (used && c < 52u) ? c : 0u;
            if (used) {
                known |= card_bit(cc);
                if (k >= 2) { bset |= card_bit(cc); bsuit += 1u << (4 * card_suit(cc)); lit_add_card(lb, cc); }
                else lit_add_card(lh, cc);
            }
        }
        ok = ok && (uint32_t)__popcll(known) == nb + 2;
        if (!ok) {
            if (lane < 3) out[idx * 3 + lane] = 0xFFFFFFFFu;
            continue;
        }
        const uint32_t ours_now = lit_category(lit_merge(lb, lh));                                 // :163
        LitHand rb = lb;                                    // the board's rank counts alone
        rb.sc = 0; rb.cards = 0; rb.dup = 0;
        // unseen cards per rank: four minus the known ones (bit 13 s + r of the card set); per suit: the complement of the suit's ranks
        auto unseen = [&](uint32_t r) { return 4u - (uint32_t)__popcll((known >> r) & 0x0000008004002001ull); };
        auto nf_of = [&](uint32_t x, uint32_t y) {
            LitHand h = rb;
            h.rc += (1ull << (4 * x)) + (1ull << (4 * y));
            h.rmask |= (1u << x) | (1u << y);
            return lit_category(h);
        };
        auto cls_of = [&](uint32_t theirs) { return ours_now > theirs ? 0u : (ours_now == theirs ? 1u : 2u); };   // :168-173
        int acc0 = 0, acc1 = 0, acc2 = 0;
        auto add = [&](uint32_t cls, int v) { acc0 += cls == 0 ? v : 0; acc1 += cls == 1 ? v : 0; acc2 += cls == 2 ? v : 0; };
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            if (ppk[p] != 0xFFu) {
                const uint32_t x = ppk[p] & 15u, y = ppk[p] >> 4;
                const uint32_t ux = unseen(x), uy = unseen(y);
                const uint32_t cnt = x != y ? ux * uy : (ux * (ux - 1)) / 2;
                if (cnt) add(cls_of(nf_of(x, y)), (int)cnt);
            }
        }
        // ---- holdings that make a flush with the board ---------------------------------------------------------------------------------
        const uint32_t f3 = (bsuit + 0x5555u) & 0x8888u;                 // the suit with >= 3 board cards (at most one)
        if (f3) {
            const uint32_t s = (uint32_t)(__ffs(f3) - 1) >> 2, bs = (bsuit >> (4 * s)) & 15u;
            const uint32_t bmask = (uint32_t)(bset >> (13 * s)) & 0x1FFFu;
            const uint32_t smask = ~(uint32_t)(known >> (13 * s)) & 0x1FFFu;          // unseen ranks of the suit
            auto flush_cat = [&](uint32_t mask, uint32_t nf) { return cyclic_run5(mask) ? 8u : ((nf == 4u || nf >= 6u) ? nf : 5u); };
            auto unseen_ns = [&](uint32_t r) { return unseen(r) - ((smask >> r) & 1u); };
            // both cards in s: pairs of distinct ranks x < y of the suit
#pragma unroll
            for (int p = 0; p < 3; ++p) {
                const uint32_t x = sqk[p] & 15u, y = sqk[p] >> 4;
                if (sqk[p] != 0xFFu && ((smask >> x) & (smask >> y) & 1u) != 0) {
                    const uint32_t nf = nf_of(x, y);
                    add(cls_of(flush_cat(bmask | (1u << x) | (1u << y), nf)), 1);
                    add(cls_of(nf), -1);
                }
            }
            if (bs >= 4) {      // one card a in s, the other outside: merged by the other card's rank
                for (uint32_t t = lane; t < 169u; t += 32) {
                    const uint32_t a = t / 13u, y = t - 13u * a;
                    const uint32_t wgt = ((smask >> a) & 1u) ? unseen_ns(y) : 0u;
                    if (wgt) {
                        const uint32_t nf = nf_of(a, y);
                        add(cls_of(flush_cat(bmask | (1u << a), nf)), (int)wgt);
                        add(cls_of(nf), -(int)wgt);
                    }
                }
            }
            if (bs == 5) {      // the board itself is the flush: holdings without a card of s, merged by rank pair
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    if (ppk[p] != 0xFFu) {
                        const uint32_t x = ppk[p] & 15u, y = ppk[p] >> 4;
                        const uint32_t wx = unseen_ns(x), wy = unseen_ns(y);
                        const uint32_t wgt = x != y ? wx * wy : (wx * (wx - 1)) / 2;
                        if (wgt) {
                            const uint32_t nf = nf_of(x, y);
                            add(cls_of(flush_cat(bmask, nf)), (int)wgt);
                            add(cls_of(nf), -(int)wgt);
                        }
                    }
                }
            }
        }
        const int a0 = __reduce_add_sync(0xFFFFFFFFu, acc0);
        const int a1 = __reduce_add_sync(0xFFFFFFFFu, acc1);
        const int a2 = __reduce_add_sync(0xFFFFFFFFu, acc2);
        if (lane < 3) out[idx * 3 + lane] = (uint32_t)(lane == 0 ? a0 : (lane == 1 ? a1 : a2));
    }
}

cudaError_t launch_hs_ref5(const uint8_t *sit, int64_t n, uint32_t *out, int sms, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    int64_t blocks = (n + 7) / 8;
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    count_launch(); hs_ref5_kernel<<<(int)blocks, 256, 0, st>>>(sit, n, out);
    return cudaGetLastError();
}

}  // namespace holdem
