// kernels_hp_turn4.cu -- turn hand strength + hand potential in Treys mode, rank-multiset formulation: the construction of
// kernels_hp_flop4.cu with ONE river card instead of a turn/river pair.  sm_100a only.
//
// What is counted (HandCalculations.py:178-226 with rank = Treys evaluate, four board cards):
//     HP[cls(P)][cmp(our7[q], R(P,q))] += 1     for every opponent holding P (1035 pairs of the 46 unseen cards) and river
//                                                card q outside P (44 per holding: 45,540 pairs per situation)
// How: R(P,q) is the suited-table rank when board + q + P holds five cards of a suit and otherwise depends on ranks only:
//     GENERIC part: for every river RANK rq and every rank pair pp: n_q(pp) * [clsr(pp)] x cmp(our7nf[rq], T[pp][rq]),
//                   13 merged rivers (+ one entry per river card that completes OUR flush) x 91 rank pairs;
//     FLUSH part:   for every (P,q) such that board + q has k >= 3 cards of a suit s and P holds >= 5-k cards of s:
//                   + [cls(P)] x cmp(our7[q], flush[board+q+P in s]) - [clsr(P)] x cmp(our7[q], T[..][..]); holdings with both
//                   cards in s one by one, holdings {a in s, c not in s} through the class histogram of a plus rank-level
//                   undo steps, holdings without a card of s (board + q already a flush) through one compare plus rank-pair
//                   undo steps; river cards outside s merged by rank (weighted lanes) unless they complete our flush.
// Execution: persistent CTA of 960 threads = fifteen groups of 64 threads, one situation each (named barriers); per
// situation one thread fetches the board's 2.9 KB row of DevTables::board4_t and the rows of own6_our7 / board4_opp6 with
// cp.async.bulk onto mbarriers; no hash lookups.  Same step descriptors and step code as the flop kernel
// (multiset_common.cuh).  Counts are bit-identical to hp_turn_treys_kernel (3-set walk) and to the plain enumeration.
#include "device_common.cuh"
#include "launch.h"
#include "multiset_common.cuh"

namespace holdem {

namespace {

using namespace ms;

constexpr int kGT = 64;              // threads per situation group
constexpr int kNG = 15;              // groups per CTA
constexpr int kThreadsT = kGT * kNG;
constexpr int kUT = 46;              // unseen cards
constexpr int kPairsT = 1035;        // C(46,2)
constexpr int kPPT = 91;
constexpr int kSegT = 64;
constexpr int kPDT = 320;            // 45 + 14 * 10 + 91 at most for the one suit with >= 3 board cards, 2 x 55 for two suits with 2
constexpr int kMaxGroupsT = 8;
constexpr int kClaimT = 4;           // situations a group claims from the global work counter at a time

struct Turn4Group {
    uint4 PD[kPDT];                    // step descriptors, as in kernels_hp_flop4.cu; w = byte offset of row pp of T
    uint16_t T[91 * 16];               // [pp][rq]: non-flush rank of board + river rank + rank pair (row of board4_t)
    uint16_t opp6nf[96];               // non-flush 6-card rank of board + rank pair (row of board4_opp6)
    uint16_t our7nf[16];               // non-flush rank of hole + board + river rank (row of own6_our7)
    unsigned long long bar[2];         // mbarriers: [0] the two small rows, [1] T
    uint16_t O7[48];                   // our final rank per river card (position)
    uint32_t ext[48];                  // river cards that give us a flush: rank | our7 << 8
    uint32_t clsw[96];                 // per rank pair: 1 << 10 clsr
    uint32_t qflush[16];               // per river rank: cards moved to ext[]
    uint32_t cnt[12];
    uint32_t bmask[4];
    uint32_t g_first[kMaxGroupsT + 1];
    uint32_t g_nq[kMaxGroupsT], g_np[kMaxGroupsT], g_nseg[kMaxGroupsT], g_seglen[kMaxGroupsT], g_pdbase[kMaxGroupsT];
    uint32_t g_sq[kMaxGroupsT];        // suit | qtype << 2 | rank-level lanes << 4
    uint32_t pd_base[4], pd_len[4];
    uint32_t smask[4];
    uint32_t hist_none[4];
    uint32_t HA[16 * 13];
    uint32_t upack[2];
    uint32_t n_groups, pd_total, next_task, ours_now, n_ext;
    uint32_t next_claim[2];
    uint32_t claim_rows[2][kClaimT];   // the rows behind the claimed work indices
    uint8_t code[48];
    uint8_t Ls[4][16];
    uint8_t Lns[4][48];
    uint8_t ns[4];
    uint8_t u[16];
    uint8_t clsr[96];
};

struct Turn4Smem {
    uint16_t flush[kFlushEntries];     // flush[fold(mask)]
    uint16_t pairs[kPairsT + 5];
    uint8_t ppxy[96];
    Turn4Group G[kNG];
};

__device__ __forceinline__ void tsync(uint32_t gid) { group_sync_n<kGT>(gid); }

}  // namespace

__global__ void __launch_bounds__(kThreadsT, 1)
hp_turn_treys_v4_kernel(const uint8_t *__restrict__ sit, int64_t n, uint32_t *__restrict__ out, DevTables T,
                        int *__restrict__ run_flag, uint32_t one, PeerOut peers, unsigned int *__restrict__ work_counter,
                        const uint32_t *__restrict__ order) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    if (run_flag != nullptr && !(*run_flag & 1)) return;   // the flop kernel saw no turn rows
    // with a work order (kernels_order.cu) only the turn-shaped rows are visited, heaviest board texture first
    const int64_t n_work = order != nullptr ? (int64_t)order[29] : n;
    Turn4Smem &S = *reinterpret_cast<Turn4Smem *>(smem_raw);
    const uint32_t gid = threadIdx.x / kGT, tid = threadIdx.x % kGT;
    const uint32_t lane = tid & 31, warp = tid >> 5;
    Turn4Group &G = S.G[gid];
    {
        for (uint32_t m = threadIdx.x; m < kFlushEntries; m += kThreadsT) S.flush[fold(m)] = __ldg(T.flush + m);
        for (uint32_t i = threadIdx.x; i < (uint32_t)kPairsT; i += kThreadsT) S.pairs[i] = __ldg(T.pairs + i);
        if (threadIdx.x < 96) {
            uint32_t y = 0;
            while (tri(y + 1) <= threadIdx.x) ++y;
            S.ppxy[threadIdx.x] = threadIdx.x < kPPT ? (uint8_t)((threadIdx.x - tri(y)) | (y << 4)) : (uint8_t)0xFF;
        }
        if (tid == 0) {
            mbar_init((uint32_t)__cvta_generic_to_shared(&G.bar[0]), 1);
            mbar_init((uint32_t)__cvta_generic_to_shared(&G.bar[1]), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    __syncthreads();
    const uint32_t bar0_s = (uint32_t)__cvta_generic_to_shared(&G.bar[0]);
    const uint32_t bar1_s = (uint32_t)__cvta_generic_to_shared(&G.bar[1]);
    const uint32_t flush_s = (uint32_t)__cvta_generic_to_shared(S.flush);
    const uint32_t t_s = (uint32_t)__cvta_generic_to_shared(G.T);
    uint32_t parity = 0;

    // situations are claimed kClaimT at a time from a global counter (see kernels_hp_flop4.cu); nullptr = static striding
    const uint32_t n_all_groups = gridDim.x * kNG, my_group = blockIdx.x * kNG + gid;
    uint32_t claim_it = 0;
    auto make_claim = [&](uint32_t slot, uint32_t it) {
        const uint32_t base = work_counter ? atomicAdd(work_counter, (unsigned)kClaimT) : (it * n_all_groups + my_group) * kClaimT;
        G.next_claim[slot] = base;
#pragma unroll
        for (uint32_t j = 0; j < (uint32_t)kClaimT; ++j) {
            const int64_t v = (int64_t)base + j;
            G.claim_rows[slot][j] = v < n_work ? (order != nullptr ? __ldg(order + kOrderHeader + n + v) : (uint32_t)v) : 0u;
        }
    };
    if (tid == 0) make_claim(0, 0);
    for (;; ++claim_it) {
    tsync(gid);
    const int64_t claim_base = G.next_claim[claim_it & 1];
    if (claim_base >= n_work) break;
    bool next_claimed = false;
    for (int64_t vidx = claim_base; vidx < n_work && vidx < claim_base + kClaimT; ++vidx) {
        const int64_t idx = G.claim_rows[claim_it & 1][vidx - claim_base];
        tsync(gid);                                                                                 // (A)
        uint2 w = __ldg(reinterpret_cast<const uint2 *>(sit) + idx);
        uint32_t cb[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) { cb[k] = (w.x >> (8 * k)) & 0xFFu; cb[4 + k] = (w.y >> (8 * k)) & 0xFFu; }
        if (!(cb[5] != 0xFFu && cb[6] == 0xFFu)) continue;              // not a four-card board: another kernel's row
        bool ok = true;
        uint64_t known = 0, bset = 0;
        uint32_t bsuit = 0, hsuit = 0;
        uint32_t rk[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
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
        ok = ok && __popcll(known) == 6;
        if (!ok) {
            if (tid == 0 && run_flag != nullptr) atomicOr(run_flag, 4);   // bit 2: a malformed row was marked
            if (tid < 16) {
                out[idx * 16 + tid] = 0xFFFFFFFFu;
#pragma unroll
                for (int k = 0; k < 8; ++k)      // constant indices: the parameter struct stays in param space
                    if (k < peers.n) peers.p[k][idx * 16 + tid] = 0xFFFFFFFFu;
            }
            continue;
        }
        if (tid < 12) G.cnt[tid] = 0;
        if (tid < 4) { G.bmask[tid] = (uint32_t)(bset >> (13 * tid)) & 0x1FFFu; G.hist_none[tid] = 0; }
        if (tid >= 16 && tid < 32) G.qflush[tid - 16] = 0;
        if (warp == 0) {
            // ---- unseen cards in (rank, suit) order; per-rank and per-suit position lists -----------------------------
            const uint32_t p0 = path_of_code4(lane), p1 = path_of_code4(lane + 32);
            const bool in0 = !((known >> p0) & 1ull);
            const bool in1 = (lane + 32 < 52) && !((known >> p1) & 1ull);
            const uint32_t m0 = __ballot_sync(0xFFFFFFFFu, in0), m1 = __ballot_sync(0xFFFFFFFFu, in1);
            const uint32_t ltm = (1u << lane) - 1u;
            if (in0) G.code[__popc(m0 & ltm)] = (uint8_t)lane;
            if (in1) G.code[__popc(m0) + __popc(m1 & ltm)] = (uint8_t)(lane + 32);
            {
                const uint32_t c4 = 4 * min(lane, 13u);
                const uint32_t cnt_r = lane < 13 ? (c4 < 32 ? __popc((m0 >> c4) & 15u) : __popc((m1 >> (c4 - 32)) & 15u)) : 0u;
                if (lane < 13) G.u[lane] = (uint8_t)cnt_r;
                const uint32_t lo = __reduce_or_sync(0xFFFFFFFFu, lane < 10 ? cnt_r << (3 * lane) : 0u);
                const uint32_t hi = __reduce_or_sync(0xFFFFFFFFu, (lane >= 10 && lane < 13) ? cnt_r << (3 * lane - 30) : 0u);
                if (lane == 0) { G.upack[0] = lo | (hi << 30); G.upack[1] = hi >> 2; }
            }
            __syncwarp();
            const uint32_t cd0 = G.code[lane], cd1 = lane + 32 < kUT ? G.code[lane + 32] : 0u;
            const bool v1 = lane + 32 < kUT;
#pragma unroll
            for (uint32_t s = 0; s < 4; ++s) {
                const bool a0 = (cd0 & 3u) == s, a1 = v1 && (cd1 & 3u) == s;
                const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, a0), b1 = __ballot_sync(0xFFFFFFFFu, a1);
                const uint32_t n0 = __popc(b0), nsu = n0 + __popc(b1);
                if (a0) G.Ls[s][__popc(b0 & ltm)] = (uint8_t)lane;
                else G.Lns[s][lane - __popc(b0 & ltm)] = (uint8_t)lane;
                if (v1) {
                    if (a1) G.Ls[s][n0 + __popc(b1 & ltm)] = (uint8_t)(lane + 32);
                    else G.Lns[s][(32 - n0) + lane - __popc(b1 & ltm)] = (uint8_t)(lane + 32);
                }
                const uint32_t rm = __reduce_or_sync(0xFFFFFFFFu, (a0 ? 1u << (cd0 >> 2) : 0u) | (a1 ? 1u << (cd1 >> 2) : 0u));
                if (lane == 0) { G.ns[s] = (uint8_t)nsu; G.smask[s] = rm; }
            }
            __syncwarp();
            {
                // ---- flush task groups and PD sections: lane = 2 * suit + (1 - qtype); qtype = cards of the river in s ------------
                const uint32_t s = min(lane >> 1, 3u), qtype = 1u - (lane & 1u);
                const uint32_t bs = (bsuit >> (4 * s)) & 15u, nsu = G.ns[s], nns = kUT - nsu;
                const uint32_t nboth = choose2u(nsu);
                const uint32_t len2 = nboth, len1 = nboth + 14u * nsu, len0 = len1 + kPPT;
                const uint32_t len = bs == 2 ? len2 : (bs == 3 ? len1 : (bs == 4 ? len0 : 0u));
                uint32_t pdo = 0;
#pragma unroll
                for (uint32_t q = 0; q < 3; ++q) {
                    const uint32_t lq = __shfl_sync(0xFFFFFFFFu, len, 2 * q);
                    if (q < s) pdo += lq;
                }
                const uint32_t k = bs + qtype, need = 5u - k;
                const uint32_t hs_s = (hsuit >> (4 * s)) & 15u;
                const bool rank_lanes = qtype == 0 && hs_s < 5;          // river cards outside s merged by rank
                const uint32_t nq = qtype == 1 ? nsu : (rank_lanes ? 13u : nns);
                const uint32_t np = need == 2 ? len2 : (need == 1 ? len1 : len0);
                const bool live = lane < 8 && bs >= 2 && k >= 3 && nq != 0 && np != 0;
                const uint32_t nseg = (np + kSegT - 1) / kSegT, seglen = live ? (np + nseg - 1) / nseg : 0u;
                const uint32_t ntask = live ? ((nq + 31) / 32) * nseg : 0u;
                const uint32_t livem = __ballot_sync(0xFFFFFFFFu, live);
                const uint32_t gi = __popc(livem & ltm);
                uint32_t first = ntask;
#pragma unroll
                for (int d = 1; d < 8; d <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, first, d);
                    if ((int)lane >= d) first += v;
                }
                if (live) {
                    G.g_first[gi] = first - ntask;
                    G.g_nq[gi] = nq; G.g_np[gi] = np; G.g_nseg[gi] = nseg; G.g_seglen[gi] = seglen;
                    G.g_pdbase[gi] = pdo;
                    G.g_sq[gi] = s | (qtype << 2) | ((uint32_t)rank_lanes << 4);
                }
                if (lane < 8 && qtype == 1) { G.pd_base[s] = pdo; G.pd_len[s] = len; }
                if (lane == 7) {
                    G.g_first[__popc(livem)] = first;
                    G.n_groups = __popc(livem);
                    G.pd_total = pdo + len;
                    G.next_task = 0;
                    G.n_ext = 0;
                }
            }
        } else {
            // ---- the situation's table rows, fetched by the TMA engine ----------------------------------------------------------
            uint32_t ours_now = 0;
            if (lane == 0) {
                uint32_t b0 = rk[2], b1 = rk[3], b2 = rk[4], b3 = rk[5];
                cswap(b0, b1); cswap(b2, b3); cswap(b0, b2); cswap(b1, b3); cswap(b1, b2);       // sort four
                const uint32_t b4 = b0 + tri(b1) + (b2 + 2) * (b2 + 1) * b2 / 6 + (b3 + 3) * (b3 + 2) * (b3 + 1) * b3 / 24;
                uint32_t a0 = rk[0], a1 = rk[1], a2 = b0, a3 = b1, a4 = b2, a5 = b3;             // six ranks: insert the hole cards
                cswap(a0, a1);
                cswap(a1, a2); cswap(a2, a3); cswap(a3, a4); cswap(a4, a5);
                cswap(a0, a1); cswap(a1, a2); cswap(a2, a3); cswap(a3, a4);
                const uint32_t m6 = a0 + tri(a1) + (a2 + 2) * (a2 + 1) * a2 / 6 + (a3 + 3) * (a3 + 2) * (a3 + 1) * a3 / 24 +
                                    (a4 + 4) * (a4 + 3) * (a4 + 2) * (a4 + 1) * a4 / 120 +
                                    (a5 + 5) * (a5 + 4) * (a5 + 3) * (a5 + 2) * (a5 + 1) * a5 / 720;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(bar0_s, 32 + 192);
                bulk_g2s((uint32_t)__cvta_generic_to_shared(G.our7nf), T.own6_our7 + (size_t)m6 * 16, 32, bar0_s);
                bulk_g2s((uint32_t)__cvta_generic_to_shared(G.opp6nf), T.board4_opp6 + (size_t)b4 * 96, 192, bar0_s);
                mbar_expect_tx(bar1_s, 91 * 16 * 2);
                bulk_g2s(t_s, T.board4_t + (size_t)b4 * (91 * 16), 91 * 16 * 2, bar1_s);
                const uint32_t fm = flush_mask(known, 5);
                ours_now = fm ? (uint32_t)S.flush[fold(fm)] : (uint32_t)__ldg(T.own6_rank6 + m6);
                G.ours_now = ours_now;
            }
            ours_now = __shfl_sync(0xFFFFFFFFu, ours_now, 0);
            mbar_wait(bar0_s, parity);
#pragma unroll
            for (uint32_t ps = 0; ps < 3; ++ps) {
                const uint32_t pp = lane + 32u * ps;
                const uint32_t v = G.opp6nf[pp];
                const uint32_t code = v == 0 ? 3u : (ours_now < v ? 0u : (ours_now == v ? 1u : 2u));
                G.clsw[pp] = v == 0 ? 0u : 1u << (10 * code);
                G.clsr[pp] = (uint8_t)code;
            }
        }
        tsync(gid);                                                                                 // (B)
        // the suit in which we can still make a flush (>= 4 of hole + board) and the one with >= 3 board cards
        const uint32_t m3 = (bsuit + 0x5555u) & 0x8888u;
        const uint32_t ms3 = m3 ? (uint32_t)(__ffs(m3) - 1) >> 2 : 4u;
        const uint32_t mb = m3 ? (bsuit >> (4 * ms3)) & 15u : 0u;                // 3 or 4 board cards of that suit
        {
            mbar_wait(bar0_s, parity);
            const uint32_t ours_now = G.ours_now;
            const uint32_t f4 = (hsuit + 0x4444u) & 0x8888u;
            const uint32_t fs = f4 ? (uint32_t)(__ffs(f4) - 1) >> 2 : 4u;
            const bool own_flush_any = f4 && ((hsuit >> (4 * fs)) & 15u) >= 5;
            const uint32_t hbmask = f4 ? (uint32_t)(known >> (13 * fs)) & 0x1FFFu : 0u;
            const uint32_t mbmask = m3 ? G.bmask[ms3] : 0u;
            // ---- our final rank per river card -----------------------------------------------------------------------------------
            if (tid < (uint32_t)kUT) {
                const uint32_t cq = G.code[tid], rq = cq >> 2, sq = cq & 3u;
                uint32_t our7 = G.our7nf[rq];
                if (own_flush_any || sq == fs) {                 // four of the suit in hand + board, and the river brings the fifth
                    our7 = S.flush[fold(hbmask | (sq == fs ? 1u << rq : 0u))];
                    G.ext[atomicAdd(&G.n_ext, 1u)] = rq | (our7 << 8);
                    atomicAdd(&G.qflush[rq], 1u);
                }
                G.O7[tid] = (uint16_t)our7;
            }
            // ---- hand strength: opponent holdings by turn class, from rank pairs; holdings that already hold a flush (both cards
            //      in a suit with three board cards, or one card in a suit with four) are moved to their true class ---------------
            for (uint32_t pp = tid; pp < (uint32_t)kPPT; pp += kGT) {
                const uint32_t xy = S.ppxy[pp], x = xy & 15u, y = xy >> 4;
                const uint32_t ux = G.u[x], uy = G.u[y];
                const uint32_t cnt = x != y ? ux * uy : (ux ? choose2u(ux) : 0u);
                const uint32_t cl = G.clsr[pp];
                if (cnt && cl < 3) atomicAdd(&G.cnt[6 + cl], cnt);
            }
            if (ms3 < 4) {
                const uint32_t nsu = G.ns[ms3], nns = kUT - nsu, nboth = choose2u(nsu);
                const uint32_t total = mb == 4 ? nboth + nsu * nns : nboth;
                for (uint32_t t = tid; t < total; t += kGT) {
                    uint32_t ra, rb, bits;
                    if (t < nboth) {
                        const uint32_t pr = S.pairs[t];
                        ra = G.code[G.Ls[ms3][pr & 0xFFu]] >> 2; rb = G.code[G.Ls[ms3][pr >> 8]] >> 2;
                        bits = (1u << ra) | (1u << rb);
                    } else {
                        const uint32_t t2 = t - nboth, a = t2 / nns, z = t2 - a * nns;
                        const uint32_t r1 = G.code[G.Ls[ms3][a]] >> 2, r2 = G.code[G.Lns[ms3][z]] >> 2;
                        ra = min(r1, r2); rb = max(r1, r2);
                        bits = 1u << r1;
                    }
                    const uint32_t opp6 = S.flush[fold(mbmask | bits)];
                    const uint32_t cls = ours_now < opp6 ? 0u : (ours_now == opp6 ? 1u : 2u);
                    atomicAdd(&G.cnt[6 + cls], 1u);
                    atomicSub(&G.cnt[6 + G.clsr[tri(rb) + ra]], 1u);
                }
            }
            // ---- step descriptors of the flush part (same layout as the flop kernel) -------------------------------------------------
            const uint32_t total = G.pd_total;
            for (uint32_t e = tid; e < total; e += kGT) {
                uint32_t s = 0;
#pragma unroll
                for (uint32_t q = 1; q < 4; ++q)
                    if (G.pd_len[q] && e >= G.pd_base[q]) s = q;
                uint32_t t = e - G.pd_base[s];
                const uint32_t nsu = G.ns[s], nboth = choose2u(nsu), sm = G.smask[s];
                const uint32_t bs = (bsuit >> (4 * s)) & 15u;
                if (t < nboth) {
                    const uint32_t pr = S.pairs[t];
                    const uint32_t ra = G.code[G.Ls[s][pr & 0xFFu]] >> 2, rb = G.code[G.Ls[s][pr >> 8]] >> 2;
                    const uint32_t pp = tri(rb) + ra, bits = (1u << ra) | (1u << rb);
                    uint32_t cls = G.clsr[pp];
                    if (bs >= 3) {                                 // the holding already has a flush on the turn
                        const uint32_t opp6 = S.flush[fold(G.bmask[s] | bits)];
                        cls = ours_now < opp6 ? 0u : (ours_now == opp6 ? 1u : 2u);
                    }
                    G.PD[e] = make_uint4((bits << 17) | (2u * fold(bits)), 1u << (10 * cls), G.clsw[pp], pp * 32u);
                } else if (t < nboth + 14u * nsu) {
                    // one card a in s (t < nboth + nsu: the add step of a, all c at once; else the undo step of (a, rank y of c)):
                    // the class of {a, c} is the flush class of a when the board has four of s, else it depends on ranks alone
                    const bool add_step = t < nboth + nsu;
                    const uint32_t t2 = t - nboth - (add_step ? 0u : nsu);
                    const uint32_t ai = add_step ? t2 : t2 / 13u, y0 = add_step ? 0u : t2 - 13u * ai, y1 = add_step ? 13u : y0 + 1u;
                    const uint32_t ra = G.code[G.Ls[s][ai]] >> 2;
                    uint32_t fcw = 0;
                    if (bs == 4) {
                        const uint32_t opp6 = S.flush[fold(G.bmask[s] | (1u << ra))];
                        fcw = 1u << (10 * (ours_now < opp6 ? 0u : (ours_now == opp6 ? 1u : 2u)));
                    }
                    uint32_t hist = 0;
                    for (uint32_t y = y0; y < y1; ++y)
                        hist += (G.u[y] - ((sm >> y) & 1u)) * (bs == 4 ? fcw : G.clsw[tri(max(ra, y)) + min(ra, y)]);
                    if (add_step) G.PD[e] = make_uint4(((2u << ra) << 16) | (2u * fold(1u << ra)), hist, 0u, 0u);
                    else {
                        const uint32_t pp = tri(max(ra, y0)) + min(ra, y0);
                        G.PD[e] = make_uint4(((2u << ra) << 16) | (2u * fold(1u << ra)), 0u,
                                             (G.u[y0] - ((sm >> y0) & 1u)) * G.clsw[pp], pp * 32u);
                        if (bs == 4) G.HA[ai * 13 + y0] = hist;
                    }
                } else {
                    const uint32_t pp = t - (nboth + 14u * nsu);
                    const uint32_t xy = S.ppxy[pp], x = xy & 15u, y = xy >> 4;
                    const uint32_t wx = G.u[x] - ((sm >> x) & 1u), wy = G.u[y] - ((sm >> y) & 1u);
                    const uint32_t wgt = x != y ? wx * wy : (wx ? choose2u(wx) : 0u);
                    G.PD[e] = make_uint4(0u, 0u, wgt * G.clsw[pp], pp * 32u);
                    if (wgt) atomicAdd(&G.hist_none[G.clsr[pp]], wgt);
                }
            }
        }
        tsync(gid);                                                                                 // (D)
        if (!next_claimed) {
            if (tid == 0) make_claim((claim_it + 1) & 1, claim_it + 1);
            next_claimed = true;
        }
        mbar_wait(bar1_s, parity);
        parity ^= 1u;

        uint32_t accL[3] = {0, 0, 0}, accG[3] = {0, 0, 0};
        const uint32_t n_groups = G.n_groups;
        const uint32_t n_flush_tasks = G.g_first[n_groups];
        const uint32_t n_entries = 13u + G.n_ext;
        const uint32_t n_gen_tasks = 2u * ((n_entries + 31) / 32);
        const uint32_t n_tasks = n_gen_tasks + n_flush_tasks;
        for (;;) {
            uint32_t task = 0;
            if (lane == 0) task = atomicAdd(&G.next_task, 1u);
            task = __shfl_sync(0xFFFFFFFFu, task, 0);
            if (task >= n_tasks) break;
            if (task >= n_gen_tasks) {
                task -= n_gen_tasks;
                // ================= flush task: 32 river lanes x a segment of steps ===========================================
                uint32_t g = 0;
                while (g + 1 < n_groups && task >= G.g_first[g + 1]) ++g;
                const uint32_t local = task - G.g_first[g];
                const uint32_t nseg = G.g_nseg[g], seglen = G.g_seglen[g], np = G.g_np[g], nq = G.g_nq[g];
                const uint32_t chunk = local / nseg, seg = local - chunk * nseg;
                const uint32_t s = G.g_sq[g] & 3u, qtype = (G.g_sq[g] >> 2) & 3u;
                const bool rank_lanes = (G.g_sq[g] >> 4) != 0;
                const uint32_t p0 = seg * seglen, p1 = min(np, p0 + seglen);
                // a chunk with few river lanes left (<= 16) is spread over the warp: each gets 32 / p2 lanes, lane stripe k of them
                // takes the steps p0 + k, p0 + k + nstr, ...
                const uint32_t live = min(32u, nq - chunk * 32);
                const uint32_t p2 = live > 16 ? 32u : (live <= 1 ? 1u : 1u << (32 - __clz(live - 1)));
                const uint32_t nstr = 32u / p2, stripe = lane / p2;
                const bool primary = stripe == 0;                 // the lane that also does the per-task extras
                const uint32_t qi = (lane & (p2 - 1)) < live ? chunk * 32 + (lane & (p2 - 1)) : nq;
                const uint32_t nsu = G.ns[s], sm = G.smask[s];
                uint32_t our = 0, fqf = 0, qb2 = 0, zpos = 0xFFu, rz = 0, qcol = 0, wl = 0, rq = 99;
                if (qi < nq) {
                    uint32_t qbits = 0;
                    if (qtype == 1) {                          // river card of s
                        const uint32_t q = G.Ls[s][qi];
                        rq = G.code[q] >> 2;
                        qbits = 1u << rq;
                        our = G.O7[q];
                        qcol = 2u * rq;
                        wl = 1;
                    } else if (rank_lanes) {                   // river rank, card outside s
                        rz = qi;
                        wl = G.u[rz] - ((sm >> rz) & 1u);
                        our = G.our7nf[rz];
                        qcol = 2u * rz;
                    } else {                                   // river card outside s, one by one (every river gives us a flush)
                        zpos = G.Lns[s][qi];
                        rz = G.code[zpos] >> 2;
                        our = G.O7[zpos];
                        qcol = 2u * rz;
                        wl = 1;
                    }
                    qb2 = qbits << 17;
                    fqf = 2u * fold(G.bmask[s] | qbits);
                }
                const uint32_t tq_s = t_s + qcol;
                const uint4 *pd = G.PD + G.g_pdbase[g] + p0;
                uint32_t aL = 0, aG = 0, sL = 0, sG = 0;
                // steps [0, nboth + n_s) have an add part (holdings with both cards in s, or one card a in s); the ones behind them
                // are rank-level undo steps: no suited-table gather, half the compares
                if (nstr > 1) {
                    pd += stripe;
                    for (uint32_t t = p0 + stripe; t < p1; t += nstr, pd += nstr) {
                        const uint4 d = *pd;
                        const uint32_t rf = lds_u16(flush_s + ((d.x & 0xFFFFu) ^ fqf));
                        const uint32_t rn = lds_u16(tq_s + d.w);
                        flush_step(aL, aG, sL, sG, our, rf, rn, d.x, qb2, d.y, d.z, one);
                    }
                } else {
                    const uint32_t split = min(p1, max(p0, choose2u(nsu) + nsu));
#pragma unroll 4
                    for (uint32_t t = p0; t < split; ++t, ++pd) {
                        const uint4 d = *pd;
                        const uint32_t rf = lds_u16(flush_s + ((d.x & 0xFFFFu) ^ fqf));
                        const uint32_t rn = lds_u16(tq_s + d.w);
                        flush_step(aL, aG, sL, sG, our, rf, rn, d.x, qb2, d.y, d.z, one);
                    }
#pragma unroll 4
                    for (uint32_t t = split; t < p1; ++t, ++pd) {
                        const uint4 d = *pd;
                        undo_step(sL, sG, our, lds_u16(tq_s + d.w), d.x, qb2, d.z, one);
                    }
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    accL[c] += (((aL >> (10 * c)) & 0x3FFu) - ((sL >> (10 * c)) & 0x3FFu)) * wl;
                    accG[c] += (((aG >> (10 * c)) & 0x3FFu) - ((sG >> (10 * c)) & 0x3FFu)) * wl;
                }
                const uint32_t need = 5u - (((bsuit >> (4 * s)) & 15u) + qtype);
                if (seg == 0 && need == 0 && qi < nq && primary) {
                    const uint32_t rf = lds_u16(flush_s + fqf);
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        if (our < rf) accL[c] += G.hist_none[c];
                        if (our > rf) accG[c] += G.hist_none[c];
                    }
                }
                if (seg == 0 && need == 1 && qtype == 0) {
                    // the river card z outside s was counted among the c of every {a, c}: take {a, z} out again (merged lanes:
                    // all wl cards of rank rz at once = the class histogram HA[a][rz] and wl undo steps of (ra, rz))
                    for (uint32_t ai = 0; ai < nsu; ++ai) {
                        const uint32_t pa = G.Ls[s][ai], ra = G.code[pa] >> 2;
                        if (wl != 0 && primary) {
                            const uint32_t rf = lds_u16(flush_s + (fqf ^ (2u * fold(1u << ra))));
                            const uint32_t pp = tri(max(ra, rz)) + min(ra, rz);
                            const uint32_t rn = lds_u16(tq_s + pp * 32u);
                            uint32_t ha, cw;
                            if (rank_lanes) {
                                ha = G.HA[ai * 13 + rz];
                                cw = wl * G.clsw[pp];
                            } else {
                                uint32_t cl = G.clsr[pp];                    // {a in s, z outside s}: a flush already when the board
                                if (mb == 4 && s == ms3) {                   // has four of s
                                    const uint32_t opp6 = lds_u16(flush_s + 2u * fold(G.bmask[s] | (1u << ra)));
                                    const uint32_t on = G.ours_now;
                                    cl = on < opp6 ? 0u : (on == opp6 ? 1u : 2u);
                                }
                                ha = 1u << (10 * cl);
                                cw = G.clsw[pp];
                            }
#pragma unroll
                            for (uint32_t c = 0; c < 3; ++c) {
                                accL[c] += (our < rn ? (cw >> (10 * c)) & 0x3FFu : 0u) - (our < rf ? (ha >> (10 * c)) & 0x3FFu : 0u);
                                accG[c] += (our > rn ? (cw >> (10 * c)) & 0x3FFu : 0u) - (our > rf ? (ha >> (10 * c)) & 0x3FFu : 0u);
                            }
                        }
                    }
                }
            } else {
                // ================= generic task: 32 merged rivers, all 91 rank pairs of P in two halves =============================
                const uint32_t e = (task >> 1) * 32 + lane;
                const bool upper = (task & 1u) != 0;
                uint32_t rq = 0, val = 0, mult = 0;
                if (e < 13u) {
                    rq = e;
                    val = G.our7nf[e];
                    mult = G.u[e] - G.qflush[e];
                } else if (e < n_entries) {
                    const uint32_t en = G.ext[e - 13u];
                    rq = en & 0xFFu;
                    val = en >> 8;
                    mult = 1;
                }
                const uint64_t m64 = (((uint64_t)G.upack[1] << 32) | G.upack[0]) - (1ull << (3 * rq));
                const uint32_t tq_s = t_s + 2u * rq;
                uint32_t lp = 0, gp = 0;
                if (!upper) {
#pragma unroll
                    for (int y = 0; y < 9; ++y) {
                        const uint32_t my = (uint32_t)(m64 >> (3 * y)) & 7u;
#pragma unroll
                        for (int x = 0; x <= y; ++x) {
                            const int pp = y * (y + 1) / 2 + x;
                            const uint32_t nn = x == y ? (my * (my - 1u)) >> 1 : ((uint32_t)(m64 >> (3 * x)) & 7u) * my;
                            cmp_add(lp, gp, val, lds_u16(tq_s + pp * 32), nn, G.clsw[pp]);
                        }
                    }
                } else {
#pragma unroll
                    for (int y = 9; y < 13; ++y) {
                        const uint32_t my = (uint32_t)(m64 >> (3 * y)) & 7u;
#pragma unroll
                        for (int x = 0; x <= y; ++x) {
                            const int pp = y * (y + 1) / 2 + x;
                            const uint32_t nn = x == y ? (my * (my - 1u)) >> 1 : ((uint32_t)(m64 >> (3 * x)) & 7u) * my;
                            cmp_add(lp, gp, val, lds_u16(tq_s + pp * 32), nn, G.clsw[pp]);
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    accL[c] += ((lp >> (10 * c)) & 0x3FFu) * mult;
                    accG[c] += ((gp >> (10 * c)) & 0x3FFu) * mult;
                }
            }
        }
#pragma unroll
        for (uint32_t c = 0; c < 3; ++c) {
            const uint32_t l = __reduce_add_sync(0xFFFFFFFFu, accL[c]);
            const uint32_t g2 = __reduce_add_sync(0xFFFFFFFFu, accG[c]);
            if (lane == 0) {
                if (l) atomicAdd(&G.cnt[c], l);
                if (g2) atomicAdd(&G.cnt[3 + c], g2);
            }
        }
        tsync(gid);                                                                                 // (E)
        if (tid < 16) {
            uint32_t v;
            if (tid < 3) v = G.cnt[6 + tid];
            else if (tid < 12) {
                const uint32_t cls = (tid - 3) / 3, o = (tid - 3) % 3;
                const uint32_t lt = G.cnt[cls], gt = G.cnt[3 + cls], all = 44u * G.cnt[6 + cls];
                v = o == 0 ? lt : (o == 2 ? gt : all - lt - gt);
            } else if (tid < 15) v = G.cnt[6 + tid - 12];
            else v = 4u;
            out[idx * 16 + tid] = v;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k < peers.n) peers.p[k][idx * 16 + tid] = v;
        }
    }
    if (!next_claimed && tid == 0) make_claim((claim_it + 1) & 1, claim_it + 1);
    }
}

cudaError_t launch_hp_turn_treys_v4(const DevTables &T, const uint8_t *sit, int64_t n, uint32_t *out, int sms,
                                    int *run_flag, const PeerOut &peers, cudaStream_t st, const uint32_t *order_scratch) {
    if (n == 0) return cudaSuccess;
    const size_t smem = sizeof(Turn4Smem);
    static_assert(sizeof(Turn4Smem) <= 227 * 1024, "fifteen situation groups must fit the 227 KB of one CTA");
    cudaError_t e = cudaFuncSetAttribute(hp_turn_treys_v4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int64_t blocks = sms;
    if (blocks * kNG > n) blocks = (n + kNG - 1) / kNG;
    // run_flag, when given, is the call's 16-byte slot: [0] flag word, [2] this kernel's work counter (zeroed with the slot)
    if (n >= ((int64_t)1 << 31)) return cudaErrorInvalidValue;
    count_launch(); hp_turn_treys_v4_kernel<<<(int)blocks, kThreadsT, smem, st>>>(sit, n, out, T, run_flag, 1u, peers,
                                                                  run_flag ? reinterpret_cast<unsigned int *>(run_flag) + 2 : nullptr,
                                                                  order_scratch);
    return cudaGetLastError();
}

// The same launch for the variant entry point: the slot only lends its work counter (word 2); the flag word is not consulted.
cudaError_t launch_hp_turn_treys_v4_counter(const DevTables &T, const uint8_t *sit, int64_t n, uint32_t *out, int sms,
                                            int *slot, const PeerOut &peers, cudaStream_t st, const uint32_t *order_scratch) {
    if (n == 0) return cudaSuccess;
    if (n >= ((int64_t)1 << 31)) return cudaErrorInvalidValue;
    const size_t smem = sizeof(Turn4Smem);
    cudaError_t e = cudaFuncSetAttribute(hp_turn_treys_v4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int64_t blocks = sms;
    if (blocks * kNG > n) blocks = (n + kNG - 1) / kNG;
    count_launch(); hp_turn_treys_v4_kernel<<<(int)blocks, kThreadsT, smem, st>>>(sit, n, out, T, nullptr, 1u, peers,
                                                                  reinterpret_cast<unsigned int *>(slot) + 2, order_scratch);
    return cudaGetLastError();
}

// Result rows -> peer buffers, for the rows that were not written there by the kernel that produced them.
__global__ void __launch_bounds__(256)
rows_to_peers_kernel(const uint8_t *__restrict__ sit, int64_t n, const uint32_t *__restrict__ out, PeerOut peers,
                     const int *__restrict__ run_flag) {
    if (run_flag != nullptr && !(*run_flag & 2)) return;
    const int64_t total = n * 16, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        if (run_flag != nullptr && sit[(i >> 4) * 8 + 6] == 0xFFu) continue;   // only five-card boards
        const uint32_t v = out[i];
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (k < peers.n) peers.p[k][i] = v;
    }
}

cudaError_t launch_rows_to_peers(const uint8_t *sit, int64_t n, const uint32_t *out, const PeerOut &peers, int sms,
                                 const int *run_flag, cudaStream_t st) {
    if (n == 0 || peers.n == 0) return cudaSuccess;
    int64_t blocks = (n * 16 + 255) / 256;
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    count_launch(); rows_to_peers_kernel<<<(int)blocks, 256, 0, st>>>(sit, n, out, peers, run_flag);
    return cudaGetLastError();
}

}  // namespace holdem
