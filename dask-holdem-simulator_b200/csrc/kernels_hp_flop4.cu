// kernels_hp_flop4.cu -- flop hand strength + hand potential in Treys mode, rank-multiset formulation ("v4").
// One CTA of 512 threads per situation, persistent grid.  sm_100a only.
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
//                        counts) and clsr = the flop class from ranks alone; Q's of equal rank pair and equal our7 are
//                        merged, so this is ~100 groups x 91 rank pairs per situation;
//                     +  FLUSH part: for every (P,Q) such that board+Q has k >= 3 cards of a suit s and P holds >= 5-k
//                        cards of s (about 8 K pairs on a rainbow flop, 40 K two-tone, 140-210 K monotone):
//                            + [cls(P)]  x cmp(our7[Q], flush[ranks of board+Q+P in s])
//                            - [clsr(P)] x cmp(our7[Q], NF(board + ranks(Q) + ranks(P)))        (undo the generic term)
//   Our own flushes need no special case: our7[Q] is evaluated per card pair Q.
//
// Shared memory per CTA (~60 KB, 3 CTAs per SM): flush[8192] u16; T[91][91] u16 = NF rank of board + rank pair q + rank
// pair p (expanded from the 1820 four-rank multisets); W[47][47] u16 = our7 | cls << 13 per unseen pair; PD[<=1081] 16-byte
// descriptors of the opponent holdings that can make a flush, per suit ordered [both cards in s | one | none].
// Flush part mapping: a warp task = 32 runouts Q of one (suit, cards-of-Q-in-suit) group x a segment of <= 64 holdings P;
// lanes own Q (its rank, suit mask, T row and position mask live in registers), the warp walks P with one broadcast
// 16-byte load per step; the per-step work is 2 gathers (flush, T), 4 compares and 4 predicated adds of packed class words.
#include "device_common.cuh"
#include "launch.h"

namespace holdem {

namespace {

constexpr int kThreads4 = 512;
constexpr int kWarps4 = kThreads4 / 32;
constexpr int kU = 47;
constexpr int kPairs4 = 1081;
constexpr int kNf4 = 1820;
constexpr int kPP = 91;            // rank pairs x <= y, index y(y+1)/2 + x
constexpr int kTPitch = 94;        // u16 units: 47 words per row (odd) so that lanes on different rows spread over the banks
constexpr int kSeg = 64;           // longest P segment of a flush task (packed 8-bit counters hold 255)
constexpr int kMaxGroups = 12;

struct Flop4Smem {
    uint4 PD[kPairs4 + 3];             // x = position mask lo, y = position mask hi | rank bits in s << 16,
                                       // z = 1 << 8 cls(P) | (2 pp) << 24, w = 1 << 8 clsr(pp)
    uint16_t flush[kFlushEntries];
    uint16_t T[kPP * kTPitch];
    uint16_t W[kU * kU + 3];           // [i*47+j], i<j: our7 | cls << 13
    uint16_t NF4[kNf4 + 4];
    uint32_t cardkey[16];              // additive key by rank
    uint32_t clsw[96];                 // per rank pair: 1 << 8 clsr, 0 when the pair cannot exist with this board
    uint32_t cnt[12];                  // lt[3], gt[3], class totals[3]
    uint32_t bmask[4];                 // board rank mask per suit
    uint32_t g_first[kMaxGroups + 1];  // first task id of each flush group; [n_groups] = total
    uint32_t g_nq[kMaxGroups], g_np[kMaxGroups], g_nseg[kMaxGroups], g_seglen[kMaxGroups], g_pdbase[kMaxGroups];
    uint32_t g_sq[kMaxGroups];         // suit | qtype << 2
    uint32_t pd_base[4], pd_len[4];    // PD section per suit
    uint32_t n_groups, pd_total, next_task;
    uint8_t code[48];                  // unseen cards, rank*4+suit, ascending
    uint8_t Ls[4][16];                 // positions of the unseen cards of suit s
    uint8_t Lns[4][48];                // positions of the unseen cards not of suit s
    uint8_t ns[4];
    uint8_t st[16];                    // first position of each rank
    uint8_t u[16];                     // unseen cards per rank
    uint8_t ppxy[96];                  // rank pair -> x | y << 4
};

__device__ __forceinline__ uint32_t path_of_code4(uint32_t code) { return (code & 3u) * 13u + (code >> 2); }
__device__ __forceinline__ uint32_t tri(uint32_t n) { return n * (n + 1) / 2; }          // C(n+1,2)
__device__ __forceinline__ uint32_t choose2u(uint32_t n) { return n * (n - 1) / 2; }    // n >= 1 or n == 0

}  // namespace

__global__ void __launch_bounds__(kThreads4)
hp_flop_treys_v4_kernel(const uint8_t *__restrict__ sit, int64_t n, uint32_t *__restrict__ out, DevTables T,
                        int *__restrict__ other_rows) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    Flop4Smem &S = *reinterpret_cast<Flop4Smem *>(smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(T.flush);
        uint4 *dst = reinterpret_cast<uint4 *>(S.flush);
        for (uint32_t i = tid; i < kFlushEntries / 8; i += kThreads4) dst[i] = src[i];
        if (tid < 16) S.cardkey[tid] = tid < 13 ? c_rank_key[tid] : 0u;
        if (tid < 96) {
            uint32_t y = 0;
            while (tri(y + 1) <= tid) ++y;
            S.ppxy[tid] = tid < kPP ? (uint8_t)((tid - tri(y)) | (y << 4)) : (uint8_t)0xFF;
        }
    }
    const NfHash h5 = T.nf[0], h7 = T.nf[2];
    int saw_other = 0;
    // lane-static rank pairs of the generic part: pass p handles pp = lane + 32 p
    uint32_t gx[3], gy[3];
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        const uint32_t pp = lane + 32u * p;
        uint32_t y = 0;
        while (tri(y + 1) <= pp) ++y;
        gx[p] = pp - tri(y);
        gy[p] = y;
    }

    for (int64_t idx = blockIdx.x; idx < n; idx += gridDim.x) {
        __syncthreads();                                                                            // (A)
        uint2 w = __ldg(reinterpret_cast<const uint2 *>(sit) + idx);
        uint32_t cb[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) { cb[k] = (w.x >> (8 * k)) & 0xFFu; cb[4 + k] = (w.y >> (8 * k)) & 0xFFu; }
        const bool flop = cb[5] == 0xFFu && cb[6] == 0xFFu;
        if (!flop) { saw_other |= (cb[6] == 0xFFu) ? 1 : 2; continue; }
        bool ok = true;
        uint64_t known = 0, bset = 0;
        uint32_t kb = 0, khole = 0, bsuit = 0, bcnt = 0;   // bsuit: 4 x 4-bit board suit counts; bcnt: 13 x 2-bit rank counts
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const uint32_t c = cb[k];
            ok = ok && c < 52u;
            const uint32_t cc = c < 52u ? c : 0u;
            known |= card_bit(cc);
            const uint32_t r = card_rank(cc);
            if (k >= 2) {
                bset |= card_bit(cc);
                kb += S.cardkey[r];
                bsuit += 1u << (4 * card_suit(cc));
                bcnt += 1u << (2 * r);
            } else khole += S.cardkey[r];
        }
        ok = ok && __popcll(known) == 5;
        if (!ok) {
            if (tid < 16) out[idx * 16 + tid] = 0xFFFFFFFFu;
            continue;
        }
        if (tid < 12) S.cnt[tid] = 0;
        if (tid < 4) S.bmask[tid] = (uint32_t)(bset >> (13 * tid)) & 0x1FFFu;
        if (warp == 0) {
            // ---- unseen cards in (rank, suit) order; per-rank and per-suit position lists -----------------------------
            const uint32_t p0 = path_of_code4(lane), p1 = path_of_code4(lane + 32);
            const bool in0 = !((known >> p0) & 1ull);
            const bool in1 = (lane + 32 < 52) && !((known >> p1) & 1ull);
            const uint32_t m0 = __ballot_sync(0xFFFFFFFFu, in0), m1 = __ballot_sync(0xFFFFFFFFu, in1);
            const uint32_t ltm = (1u << lane) - 1u;
            if (in0) S.code[__popc(m0 & ltm)] = (uint8_t)lane;
            if (in1) S.code[__popc(m0) + __popc(m1 & ltm)] = (uint8_t)(lane + 32);
            if (lane < 14) {   // rank r owns codes 4r..4r+3: start position = unseen codes below 4r
                const uint32_t c4 = 4 * lane;
                const uint32_t below = c4 < 32 ? __popc(m0 & ((1u << c4) - 1u))
                                               : __popc(m0) + __popc(m1 & ((1u << (c4 - 32)) - 1u));
                S.st[lane] = (uint8_t)below;
                if (lane < 13) {
                    const uint32_t cnt_r = c4 < 32 ? __popc((m0 >> c4) & 15u) : __popc((m1 >> (c4 - 32)) & 15u);
                    S.u[lane] = (uint8_t)cnt_r;
                }
            }
            __syncwarp();
            // positions lane and lane+32 by suit
            const uint32_t cd0 = S.code[lane], cd1 = lane + 32 < kU ? S.code[lane + 32] : 0u;
            const bool v1 = lane + 32 < kU;
#pragma unroll
            for (uint32_t s = 0; s < 4; ++s) {
                const bool a0 = (cd0 & 3u) == s, a1 = v1 && (cd1 & 3u) == s;
                const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, a0), b1 = __ballot_sync(0xFFFFFFFFu, a1);
                const uint32_t n0 = __popc(b0), ns = n0 + __popc(b1);
                if (a0) S.Ls[s][__popc(b0 & ltm)] = (uint8_t)lane;
                else S.Lns[s][lane - __popc(b0 & ltm)] = (uint8_t)lane;
                if (v1) {
                    if (a1) S.Ls[s][n0 + __popc(b1 & ltm)] = (uint8_t)(lane + 32);
                    else S.Lns[s][(32 - n0) + lane - __popc(b1 & ltm)] = (uint8_t)(lane + 32);
                }
                if (lane == 0) S.ns[s] = (uint8_t)ns;
            }
            __syncwarp();
            if (lane == 0) {
                // ---- flush task groups and PD sections --------------------------------------------------------------
                uint32_t ng = 0, first = 0, pdo = 0;
                for (uint32_t s = 0; s < 4; ++s) {
                    const uint32_t bs = (bsuit >> (4 * s)) & 15u, nsu = S.ns[s], nns = kU - nsu;
                    const uint32_t nboth = nsu ? choose2u(nsu) : 0u, none = nsu * nns;
                    uint32_t len = 0;
                    if (bs == 1) len = nboth;
                    else if (bs == 2) len = nboth + none;
                    else if (bs == 3) len = kPairs4;
                    S.pd_base[s] = pdo;
                    S.pd_len[s] = len;
                    if (bs == 0) continue;
                    for (int qtype = 2; qtype >= 0; --qtype) {
                        const uint32_t k = bs + (uint32_t)qtype;
                        if (k < 3) continue;
                        const uint32_t need = 5 - k;   // 2, 1 or 0 cards of s in P
                        const uint32_t nq = qtype == 2 ? nboth : (qtype == 1 ? none : (nns ? choose2u(nns) : 0u));
                        const uint32_t np = need == 2 ? nboth : (need == 1 ? nboth + none : (uint32_t)kPairs4);
                        if (nq == 0 || np == 0) continue;
                        const uint32_t nseg = (np + kSeg - 1) / kSeg, seglen = (np + nseg - 1) / nseg;
                        S.g_first[ng] = first;
                        S.g_nq[ng] = nq; S.g_np[ng] = np; S.g_nseg[ng] = nseg; S.g_seglen[ng] = seglen;
                        S.g_pdbase[ng] = pdo;
                        S.g_sq[ng] = s | ((uint32_t)qtype << 2);
                        first += ((nq + 31) / 32) * nseg;
                        ++ng;
                    }
                    pdo += len;
                }
                S.g_first[ng] = first;
                S.n_groups = ng;
                S.pd_total = pdo;
                S.next_task = 0;
            }
        } else {
            // ---- NF4[1820]: non-flush rank of board + every multiset of four ranks -------------------------------------
            for (uint32_t e = tid - 32; e < (uint32_t)kNf4; e += kThreads4 - 32) {
                const uint32_t rr = __ldg(T.nf4_ranks + e);
                uint32_t key = kb;
                bool valid = true;
                uint32_t prev = 99, run = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t r = (rr >> (4 * q)) & 15u;
                    key += S.cardkey[r];
                    run = (r == prev) ? run + 1 : 1;
                    prev = r;
                    valid = valid && (run + ((bcnt >> (2 * r)) & 3u) <= 4);
                }
                S.NF4[e] = valid ? (uint16_t)nf_lookup(h7, key) : (uint16_t)0;
            }
        }
        __syncthreads();                                                                            // (B)
        {
            // ---- W: our final rank with the pair as runout, flop class of the pair as a holding ------------------------------
            const uint32_t ours_now = treys_rank(h5, S.flush, kb + khole, known);
            uint32_t c0 = 0, c1 = 0, c2 = 0;
            for (uint32_t p = tid; p < (uint32_t)kPairs4; p += kThreads4) {
                const uint32_t pr = __ldg(T.pairs + p);
                const uint32_t i = pr & 0xFFu, j = pr >> 8;
                const uint32_t ci = S.code[i], cj = S.code[j];
                const uint32_t pi = path_of_code4(ci), pj = path_of_code4(cj);
                const uint32_t key = S.cardkey[ci >> 2] + S.cardkey[cj >> 2];
                const uint64_t cs = card_bit(pi) | card_bit(pj);
                const uint32_t our7 = treys_rank(h7, S.flush, kb + khole + key, known | cs);
                const uint32_t opp5 = treys_rank(h5, S.flush, kb + key, bset | cs);
                const uint32_t cls = ours_now < opp5 ? 0u : (ours_now == opp5 ? 1u : 2u);
                c0 += cls == 0; c1 += cls == 1; c2 += cls == 2;
                S.W[i * kU + j] = (uint16_t)(our7 | (cls << 13));
            }
            c0 = __reduce_add_sync(0xFFFFFFFFu, c0);
            c1 = __reduce_add_sync(0xFFFFFFFFu, c1);
            c2 = __reduce_add_sync(0xFFFFFFFFu, c2);
            if (lane == 0) { atomicAdd(&S.cnt[6], c0); atomicAdd(&S.cnt[7], c1); atomicAdd(&S.cnt[8], c2); }
            // ---- rank-level flop class per rank pair -------------------------------------------------------------------------
            if (tid >= kThreads4 - 96) {
                const uint32_t pp = tid - (kThreads4 - 96);
                uint32_t word = 0;
                if (pp < (uint32_t)kPP) {
                    const uint32_t xy = S.ppxy[pp], x = xy & 15u, y = xy >> 4;
                    const uint32_t cx = ((bcnt >> (2 * x)) & 3u) + 1 + (x == y), cy = ((bcnt >> (2 * y)) & 3u) + 1;
                    if (cx <= 4 && cy <= 4) {
                        const uint32_t v = nf_lookup(h5, kb + S.cardkey[x] + S.cardkey[y]);
                        word = 1u << (8 * (ours_now < v ? 0u : (ours_now == v ? 1u : 2u)));
                    }
                }
                S.clsw[pp] = word;
            }
            // ---- T[qp][pp] = NF4[multiset(qp + pp)] --------------------------------------------------------------------------
            for (uint32_t e = tid; e < (uint32_t)(kPP * kPP); e += kThreads4) {
                const uint32_t qp = e / kPP, pp = e - qp * kPP;
                S.T[qp * kTPitch + pp] = S.NF4[__ldg(T.idx4 + e)];
            }
        }
        __syncthreads();                                                                            // (C)
        {
            // ---- P descriptors ---------------------------------------------------------------------------------------------------
            const uint32_t total = S.pd_total;
            for (uint32_t e = tid; e < total; e += kThreads4) {
                uint32_t s = 0;
#pragma unroll
                for (uint32_t q = 1; q < 4; ++q)
                    if (S.pd_len[q] && e >= S.pd_base[q]) s = q;
                uint32_t t = e - S.pd_base[s];
                const uint32_t nsu = S.ns[s], nns = kU - nsu, nboth = nsu ? choose2u(nsu) : 0u;
                uint32_t pa, pb;
                if (t < nboth) {
                    const uint32_t pr = __ldg(T.pairs + t);
                    pa = S.Ls[s][pr & 0xFFu]; pb = S.Ls[s][pr >> 8];
                } else if (t < nboth + nsu * nns) {
                    t -= nboth;
                    const uint32_t a = t / nns, z = t - a * nns;
                    const uint32_t x = S.Ls[s][a], y = S.Lns[s][z];
                    pa = min(x, y); pb = max(x, y);
                } else {
                    t -= nboth + nsu * nns;
                    const uint32_t pr = __ldg(T.pairs + t);
                    pa = S.Lns[s][pr & 0xFFu]; pb = S.Lns[s][pr >> 8];
                }
                const uint32_t ca = S.code[pa], cbb = S.code[pb];
                const uint32_t ra = ca >> 2, rb = cbb >> 2;
                const uint32_t cls = S.W[pa * kU + pb] >> 13;
                const uint32_t pp = tri(rb) + ra;
                const uint32_t pbits = ((ca & 3u) == s ? 1u << ra : 0u) | ((cbb & 3u) == s ? 1u << rb : 0u);
                const uint64_t pm = (1ull << pa) | (1ull << pb);
                S.PD[e] = make_uint4((uint32_t)pm, (uint32_t)(pm >> 32) | (pbits << 16),
                                     (1u << (8 * cls)) | ((2u * pp) << 24), S.clsw[pp]);
            }
        }
        // per-situation lane constants of the generic part
        uint32_t gux[3], guy[3], gcls[3];
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            const uint32_t pp = lane + 32u * p;
            gux[p] = pp < (uint32_t)kPP ? S.u[gx[p]] : 0u;
            guy[p] = pp < (uint32_t)kPP ? S.u[gy[p]] : 0u;
            const uint32_t cw = S.clsw[pp];
            gcls[p] = cw == 1u ? 0u : (cw == 0x100u ? 1u : (cw == 0x10000u ? 2u : 3u));
        }
        __syncthreads();                                                                            // (D)

        uint32_t gl[3] = {0, 0, 0}, gg[3] = {0, 0, 0};     // generic part: per pass (class = gcls[p])
        uint32_t fl[3] = {0, 0, 0}, fg[3] = {0, 0, 0};     // flush part: net per class (add - undo), wraps mod 2^32
        const uint32_t n_groups = S.n_groups;
        const uint32_t n_flush_tasks = S.g_first[n_groups];
        const uint32_t n_tasks = n_flush_tasks + kPP;
        for (;;) {
            uint32_t task = 0;
            if (lane == 0) task = atomicAdd(&S.next_task, 1u);
            task = __shfl_sync(0xFFFFFFFFu, task, 0);
            if (task >= n_tasks) break;
            if (task < n_flush_tasks) {
                // ================= flush task: 32 runouts Q x a segment of holdings P ==================================
                uint32_t g = 0;
                while (g + 1 < n_groups && task >= S.g_first[g + 1]) ++g;
                const uint32_t local = task - S.g_first[g];
                const uint32_t nseg = S.g_nseg[g], seglen = S.g_seglen[g], np = S.g_np[g], nq = S.g_nq[g];
                const uint32_t chunk = local / nseg, seg = local - chunk * nseg;
                const uint32_t s = S.g_sq[g] & 3u, qtype = S.g_sq[g] >> 2;
                const uint32_t p0 = seg * seglen, p1 = min(np, p0 + seglen);
                const uint32_t qi = chunk * 32 + lane;
                uint32_t our = 0, fq = 0, qlo = 0xFFFFFFFFu, qhi = 0x7FFFu;
                const char *qrow = reinterpret_cast<const char *>(S.T);
                if (qi < nq) {
                    const uint32_t nsu = S.ns[s], nns = kU - nsu;
                    uint32_t i, j;
                    if (qtype == 1) {
                        const uint32_t a = qi / nns, z = qi - a * nns;
                        const uint32_t x = S.Ls[s][a], y = S.Lns[s][z];
                        i = min(x, y); j = max(x, y);
                    } else {
                        const uint32_t pr = __ldg(T.pairs + qi);
                        const uint8_t *L = qtype == 2 ? S.Ls[s] : S.Lns[s];
                        i = L[pr & 0xFFu]; j = L[pr >> 8];
                    }
                    const uint32_t ci = S.code[i], cj = S.code[j];
                    const uint32_t ri = ci >> 2, rj = cj >> 2;
                    our = S.W[i * kU + j] & 0x1FFFu;
                    fq = S.bmask[s] | ((ci & 3u) == s ? 1u << ri : 0u) | ((cj & 3u) == s ? 1u << rj : 0u);
                    const uint64_t qm = (1ull << i) | (1ull << j);
                    qlo = (uint32_t)qm; qhi = (uint32_t)(qm >> 32);
                    qrow += (tri(rj) + ri) * (kTPitch * 2);
                }
                const uint4 *pd = S.PD + S.g_pdbase[g] + p0;
                uint32_t aL = 0, aG = 0, sL = 0, sG = 0;
#pragma unroll 4
                for (uint32_t t = p0; t < p1; ++t, ++pd) {
                    const uint4 d = *pd;
                    const bool valid = ((d.x & qlo) | (d.y & qhi)) == 0u;
                    const uint32_t rf = S.flush[fq | (d.y >> 16)];
                    const uint32_t rn = *reinterpret_cast<const uint16_t *>(qrow + (d.z >> 24));
                    if (valid && our < rf) aL += d.z;
                    if (valid && our > rf) aG += d.z;
                    if (valid && our < rn) sL += d.w;
                    if (valid && our > rn) sG += d.w;
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    fl[c] += ((aL >> (8 * c)) & 0xFFu) - ((sL >> (8 * c)) & 0xFFu);
                    fg[c] += ((aG >> (8 * c)) & 0xFFu) - ((sG >> (8 * c)) & 0xFFu);
                }
            } else {
                // ================= generic task: one rank pair of Q, lanes over the rank pairs of P ========================
                const uint32_t qp = task - n_flush_tasks;
                const uint32_t qxy = S.ppxy[qp], q1 = qxy & 15u, q2 = qxy >> 4;
                const uint32_t u1 = S.u[q1], u2 = S.u[q2];
                const uint32_t nq = q1 != q2 ? u1 * u2 : (u1 ? choose2u(u1) : 0u);
                if (nq == 0) continue;
                uint32_t val = 0xFFFFFFFFu;
                if (lane < nq) {
                    uint32_t a, b;
                    if (q1 != q2) {
                        a = u2 == 1 ? lane : (u2 == 2 ? lane >> 1 : (u2 == 4 ? lane >> 2 : (lane * 11u) >> 5));
                        b = lane - a * u2;
                    } else {
                        a = (0x210100u >> (4 * lane)) & 15u;     // pairs of {0..3} in colex order: 01 02 12 03 13 23
                        b = (0x333221u >> (4 * lane)) & 15u;
                    }
                    const uint32_t i = S.st[q1] + a, j = S.st[q2] + b;
                    val = S.W[i * kU + j] & 0x1FFFu;
                }
                const uint16_t *row = S.T + qp * kTPitch;
                uint32_t R[3], nn[3];
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    const uint32_t pp = lane + 32u * p;
                    R[p] = pp < (uint32_t)kPP ? row[pp] : 0u;
                    const uint32_t mx = gux[p] - (gx[p] == q1) - (gx[p] == q2);
                    const uint32_t my = guy[p] - (gy[p] == q1) - (gy[p] == q2);
                    nn[p] = pp < (uint32_t)kPP ? (gx[p] != gy[p] ? mx * my : (mx ? choose2u(mx) : 0u)) : 0u;
                }
                uint32_t remaining = __ballot_sync(0xFFFFFFFFu, lane < nq);
                while (remaining) {
                    const uint32_t v = __shfl_sync(0xFFFFFFFFu, val, __ffs(remaining) - 1);
                    const uint32_t same = __ballot_sync(0xFFFFFFFFu, val == v) & remaining;
                    const uint32_t gmul = __popc(same);
                    remaining ^= same;
#pragma unroll
                    for (int p = 0; p < 3; ++p) {
                        const uint32_t wgt = nn[p] * gmul;
                        if (v < R[p]) gl[p] += wgt;
                        if (v > R[p]) gg[p] += wgt;
                    }
                }
            }
        }
        // ---- reduce: per class over the warp, then into the CTA counters -------------------------------------------------------
#pragma unroll
        for (uint32_t c = 0; c < 3; ++c) {
            uint32_t l = fl[c], g2 = fg[c];
#pragma unroll
            for (int p = 0; p < 3; ++p) {
                l += gcls[p] == c ? gl[p] : 0u;
                g2 += gcls[p] == c ? gg[p] : 0u;
            }
            l = __reduce_add_sync(0xFFFFFFFFu, l);
            g2 = __reduce_add_sync(0xFFFFFFFFu, g2);
            if (lane == 0) {
                if (l) atomicAdd(&S.cnt[c], l);
                if (g2) atomicAdd(&S.cnt[3 + c], g2);
            }
        }
        __syncthreads();                                                                            // (E)
        if (tid < 16) {
            uint32_t v;
            if (tid < 3) v = S.cnt[6 + tid];
            else if (tid < 12) {
                const uint32_t cls = (tid - 3) / 3, o = (tid - 3) % 3;
                const uint32_t lt = S.cnt[cls], gt = S.cnt[3 + cls], all = 990u * S.cnt[6 + cls];
                v = o == 0 ? lt : (o == 2 ? gt : all - lt - gt);
            } else if (tid < 15) v = S.cnt[6 + tid - 12];
            else v = 3u;
            out[idx * 16 + tid] = v;
        }
    }
    if (saw_other && tid == 0) atomicOr(other_rows, saw_other);
}

cudaError_t launch_hp_flop_treys_v4(const DevTables &T, const uint8_t *sit, int64_t n, uint32_t *out, int sms,
                                    int *other_rows_flag, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const size_t smem = sizeof(Flop4Smem);
    cudaError_t e = cudaFuncSetAttribute(hp_flop_treys_v4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    static int bps = 0;
    if (bps == 0) {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, hp_flop_treys_v4_kernel, kThreads4, smem);
        if (e != cudaSuccess) return e;
        if (bps < 1) bps = 1;
    }
    int64_t blocks = (int64_t)sms * bps;
    if (blocks > n) blocks = n;
    e = cudaMemsetAsync(other_rows_flag, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    hp_flop_treys_v4_kernel<<<(int)blocks, kThreads4, smem, st>>>(sit, n, out, T, other_rows_flag);
    return cudaGetLastError();
}

}  // namespace holdem
