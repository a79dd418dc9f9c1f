// kernels_hp_flop.cu -- the hot kernel: flop hand strength + hand potential in Treys mode (HandCalculations.py:178-226
// with rank = Treys evaluate), one CTA per situation.  sm_100a only.
//
// Algorithm (what is and is not evaluated; SURVEY.md section 8d asks for this to be stated):
//   The reference loop visits every (opponent holding P, runout Q) pair: 1081 x 990 = 1,070,190 pairs per flop
//   situation, and evaluates our 7-card hand and the opponent's 7-card hand for each.  The opponent's final hand is
//   board + P + Q: it depends only on the 4-SET S = P u Q of unseen cards, and each of the C(47,4) = 178,365 4-sets
//   is produced by exactly 6 (P, Q) splits.  Our final hand depends only on Q (1,081 values).  So per situation this
//   kernel performs 178,365 opponent evaluations + 1,081 own 7-card evaluations + 1,082 five-card evaluations, and
//   for every 4-set applies its 6 splits as compare-and-count steps:
//         HP[class(P)][cmp(our7[Q], R(S))] += 1          for the 6 ways to write S = P u Q.
//   The counts are identical to the plain enumeration (hp_direct_kernel; cross-checked in tests/test_gpu_hshp.py).
//
// Data layout (shared memory, per CTA ~25 KB + 16 KB flush table):
//   unseen cards are ordered by (rank, suit), so any a<b<c<d has non-decreasing ranks and the rank multiset of the
//   four added cards has the dense index  C(ra,1) + C(rb+1,2) + C(rc+2,3) + C(rd+3,4)  in [0,1820): NF[1820] holds
//   the non-flush 7-card rank of board + multiset (built per situation from the global perfect-hash table), so the
//   per-4-set evaluation is one add and one 16-bit shared-memory gather; flushes go through the 8192-entry suited
//   table.  W[i][j] (47 x 48 words) holds for every unseen pair  our7 << 19 | 1 << (6 * class): the high 13 bits
//   compare against R << 19 without unpacking, the low 18 bits are three 6-bit one-hot class counters that are
//   added into packed accumulators under the compare predicate.
//   One warp owns a pair (a,b); its lanes stride over the pairs (c,d) with c > b, which sit contiguously at the
//   front of a list ordered by c descending, so the list reads are conflict-free and the W row reads are
//   broadcast/stride-1.
#include "device_common.cuh"
#include "launch.h"

namespace holdem {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kUnseen = 47;
constexpr int kPairs = 1081;       // C(47,2)
constexpr int kPitch = 48;         // W row pitch (words)
constexpr int kNf = 1820;          // multisets of 4 ranks out of 13

struct FlopSmem {
    uint16_t flush[kFlushEntries];     // suited table (copied once per CTA)
    uint32_t W[kUnseen * kPitch];      // [i][j], i<j: our7 << 19 | one-hot class
    uint32_t Lw[kPairs + 3];           // pair list (c descending): W[c][d]
    uint32_t Lk[kPairs + 3];           // pair list: C3[rc] + C4[rd] | (suit nibbles of c,d) << 16
    uint16_t Lcd[kPairs + 3];          // pair list: c | d << 8
    uint16_t NF[kNf + 4];              // non-flush rank of board + 4-rank multiset
    uint32_t cardkey[64];              // additive key by path card number
    uint32_t cnt[12];                  // lt[3], gt[3] (by class), class totals[3], spare
    uint8_t code[64];                  // unseen cards, rank-major code rank*4+suit, ascending
};

__device__ __forceinline__ uint32_t path_of_code(uint32_t code) { return (code & 3u) * 13u + (code >> 2); }

__device__ __forceinline__ uint32_t choose2(uint32_t n) { return n * (n - 1) / 2; }
__device__ __forceinline__ uint32_t choose3(uint32_t n) { return n * (n - 1) * (n - 2) / 6; }
__device__ __forceinline__ uint32_t choose4(uint32_t n) { return n * (n - 1) * (n - 2) * (n - 3) / 24; }

// One (a,b) task: lanes stride over the n_cd pairs (c,d), c > b.  kFlush: some suit already holds >= 3 of the five
// cards board+a+b, so a flush is possible and the suited path is compiled in; otherwise it is skipped entirely.
template <bool kFlush>
__device__ __forceinline__ void run_task(const FlopSmem &S, uint32_t lane, uint32_t n_cd, const uint32_t *rowA,
                                         const uint32_t *rowB, uint32_t wab, uint32_t nf_ab, uint32_t sc5,
                                         uint64_t cs5, uint32_t (&acc)[6]) {
    for (uint32_t i = lane; i < n_cd; i += 32) {
        const uint32_t cd = S.Lcd[i];
        const uint32_t c = cd & 0xFFu, d = cd >> 8;
        const uint32_t wcd = S.Lw[i];
        const uint32_t k = S.Lk[i];
        const uint32_t wac = rowA[c], wad = rowA[d], wbc = rowB[c], wbd = rowB[d];
        uint32_t R = S.NF[nf_ab + (k & 0xFFFFu)];
        if (kFlush) {
            const uint32_t fl = (sc5 + (k >> 16)) & 0x8888u;   // nibbles are biased by 3: bit 3 set <=> count >= 5
            if (fl) {
                const uint32_t s = (uint32_t)(__ffs(fl) - 1) >> 2;
                const uint64_t cs = cs5 | card_bit(path_of_code(S.code[c])) | card_bit(path_of_code(S.code[d]));
                R = S.flush[(uint32_t)(cs >> (13 * s)) & 0x1FFFu];
            }
        }
        const uint32_t lo = R << 19, hi = lo + (1u << 19);
        // the six (P,Q) splits of {a,b,c,d}; "our7[Q] < R" = we end ahead, ">= R+1" = we end behind
        if (wcd < lo) acc[0] += wab;  if (wcd >= hi) acc[3] += wab;     // P=ab Q=cd
        if (wab < lo) acc[0] += wcd;  if (wab >= hi) acc[3] += wcd;     // P=cd Q=ab
        if (wbd < lo) acc[1] += wac;  if (wbd >= hi) acc[4] += wac;     // P=ac Q=bd
        if (wac < lo) acc[1] += wbd;  if (wac >= hi) acc[4] += wbd;     // P=bd Q=ac
        if (wbc < lo) acc[2] += wad;  if (wbc >= hi) acc[5] += wad;     // P=ad Q=bc
        if (wad < lo) acc[2] += wbc;  if (wad >= hi) acc[5] += wbc;     // P=bc Q=ad
    }
}

}  // namespace

__global__ void __launch_bounds__(kThreads)
hp_flop_treys_kernel(const uint8_t *__restrict__ sit, int64_t n, uint32_t *__restrict__ out, DevTables T,
                     int *__restrict__ other_rows) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    FlopSmem &S = *reinterpret_cast<FlopSmem *>(smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(T.flush);
        uint4 *dst = reinterpret_cast<uint4 *>(S.flush);
        for (uint32_t i = tid; i < kFlushEntries / 8; i += kThreads) dst[i] = src[i];
        if (tid < 64) S.cardkey[tid] = tid < 52 ? c_rank_key[tid % 13] : 0u;
    }
    const NfHash h5 = T.nf[0], h7 = T.nf[2];
    int saw_other = 0;

    for (int64_t idx = blockIdx.x; idx < n; idx += gridDim.x) {
        __syncthreads();
        // ---- parse the row (every thread, redundantly) -----------------------------------------------------------
        uint2 w = __ldg(reinterpret_cast<const uint2 *>(sit) + idx);
        uint32_t cb[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) { cb[k] = (w.x >> (8 * k)) & 0xFFu; cb[4 + k] = (w.y >> (8 * k)) & 0xFFu; }
        const bool flop = cb[5] == 0xFFu && cb[6] == 0xFFu;
        if (!flop) { saw_other |= (cb[6] == 0xFFu) ? 1 : 2; continue; }   // turn rows: hp_turn_treys_kernel; river: plain enumeration
        bool ok = true;
        uint64_t known = 0, bset = 0;
        uint32_t kb = 0, khole = 0, scb = 0x3333u;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const uint32_t c = cb[k];
            ok = ok && c < 52u;
            const uint32_t cc = c < 52u ? c : 0u;
            known |= card_bit(cc);
            if (k >= 2) { bset |= card_bit(cc); kb += S.cardkey[cc]; scb += 1u << (4 * card_suit(cc)); }
            else khole += S.cardkey[cc];
        }
        ok = ok && __popcll(known) == 5;
        if (!ok) {
            if (tid < 16) out[idx * 16 + tid] = 0xFFFFFFFFu;
            continue;
        }
        if (tid < 12) S.cnt[tid] = 0;
        // ---- unseen cards in (rank, suit) order: warp 0, lane l owns codes l and l+32 ---------------------------------
        if (warp == 0) {
            const uint32_t p0 = path_of_code(lane), p1 = path_of_code(lane + 32);
            const bool in0 = !((known >> p0) & 1ull);
            const bool in1 = (lane + 32 < 52) && !((known >> p1) & 1ull);
            const uint32_t m0 = __ballot_sync(0xFFFFFFFFu, in0), m1 = __ballot_sync(0xFFFFFFFFu, in1);
            const uint32_t lt = (1u << lane) - 1u;
            if (in0) S.code[__popc(m0 & lt)] = (uint8_t)lane;
            if (in1) S.code[__popc(m0) + __popc(m1 & lt)] = (uint8_t)(lane + 32);
        }
        // ---- NF[1820]: board + every multiset of four ranks -------------------------------------------------------------
        {
            uint32_t bcnt = 0;  // 13 x 2-bit board rank multiplicities (a flop holds at most 3 of a rank)
#pragma unroll
            for (int k = 2; k < 5; ++k) bcnt += 1u << (2 * card_rank(cb[k]));
            for (uint32_t e = tid; e < (uint32_t)kNf; e += kThreads) {
                // unrank e = C(x0,1) + C(x1,2) + C(x2,3) + C(x3,4), x0<x1<x2<x3; ranks r_i = x_i - i
                uint32_t rem = e, x3 = 3, x2 = 2, x1 = 1;
                while (choose4(x3 + 1) <= rem) ++x3;
                rem -= choose4(x3);
                while (choose3(x2 + 1) <= rem) ++x2;
                rem -= choose3(x2);
                while (choose2(x1 + 1) <= rem) ++x1;
                rem -= choose2(x1);
                const uint32_t r[4] = {rem, x1 - 1, x2 - 2, x3 - 3};
                uint32_t key = kb;
                uint64_t cnt4 = 0;  // 13 x 4-bit multiplicities of the added ranks plus the board's
#pragma unroll
                for (int q = 0; q < 4; ++q) { key += S.cardkey[r[q]]; cnt4 += 1ull << (4 * r[q]); }
                bool valid = true;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    valid = valid && (((cnt4 >> (4 * r[q])) & 15ull) + ((bcnt >> (2 * r[q])) & 3u) <= 4);
                S.NF[e] = valid ? (uint16_t)nf_lookup(h7, key) : (uint16_t)0;
            }
        }
        __syncthreads();
        // ---- per-pair words: our final rank with the pair as runout, and the pair's class as an opponent holding ---------
        {
            const uint32_t ours_now = treys_rank(h5, S.flush, kb + khole, known);
            uint32_t c0 = 0, c1 = 0, c2 = 0;
            for (uint32_t p = tid; p < (uint32_t)kPairs; p += kThreads) {
                const uint32_t pr = __ldg(T.pairs + p);
                const uint32_t i = pr & 0xFFu, j = pr >> 8;
                const uint32_t ci = S.code[i], cj = S.code[j];
                const uint32_t pi = path_of_code(ci), pj = path_of_code(cj);
                const uint32_t key = S.cardkey[pi] + S.cardkey[pj];
                const uint64_t cs = card_bit(pi) | card_bit(pj);
                const uint32_t our7 = treys_rank(h7, S.flush, kb + khole + key, known | cs);
                const uint32_t opp5 = treys_rank(h5, S.flush, kb + key, bset | cs);
                const uint32_t cls = ours_now < opp5 ? 0u : (ours_now == opp5 ? 1u : 2u);
                c0 += cls == 0; c1 += cls == 1; c2 += cls == 2;
                const uint32_t word = (our7 << 19) | (1u << (6 * cls));
                S.W[i * kPitch + j] = word;
                const uint32_t pos = choose2(46 - i) + (j - i - 1);   // list ordered by first index descending
                const uint32_t ri = ci >> 2, rj = cj >> 2;
                S.Lw[pos] = word;
                S.Lcd[pos] = (uint16_t)(i | (j << 8));
                S.Lk[pos] = (choose3(ri + 2) + choose4(rj + 3)) | (((1u << (4 * (ci & 3u))) + (1u << (4 * (cj & 3u)))) << 16);
            }
            c0 = __reduce_add_sync(0xFFFFFFFFu, c0);
            c1 = __reduce_add_sync(0xFFFFFFFFu, c1);
            c2 = __reduce_add_sync(0xFFFFFFFFu, c2);
            if (lane == 0) { atomicAdd(&S.cnt[6], c0); atomicAdd(&S.cnt[7], c1); atomicAdd(&S.cnt[8], c2); }
        }
        __syncthreads();
        // ---- main loop: one warp per (a,b), lanes over (c,d) ------------------------------------------------------------------
        uint32_t acc[6] = {0, 0, 0, 0, 0, 0};      // packed 3 x 6-bit class counters: lt x3 partitions, gt x3 partitions
        uint32_t wide[6] = {0, 0, 0, 0, 0, 0};     // lt[class], gt[class]
        uint32_t since = 0;                        // iterations since the packed counters were last drained
        for (uint32_t t = warp; t < (uint32_t)kPairs; t += kWarps) {
            const uint32_t pr = __ldg(T.pairs + t);
            const uint32_t a = pr & 0xFFu, b = pr >> 8;
            const uint32_t n_cd = choose2(46 - b);
            const uint32_t iters = (n_cd + 31) >> 5;
            if (iters == 0) continue;
            if (since + iters > 31) {   // each packed field grows by at most 2 per iteration: drain before 63
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    wide[0] += acc[q] & 63u; wide[1] += (acc[q] >> 6) & 63u; wide[2] += (acc[q] >> 12) & 63u;
                    wide[3] += acc[3 + q] & 63u; wide[4] += (acc[3 + q] >> 6) & 63u; wide[5] += (acc[3 + q] >> 12) & 63u;
                    acc[q] = 0; acc[3 + q] = 0;
                }
                since = 0;
            }
            since += iters;
            const uint32_t ca = S.code[a], cbb = S.code[b];
            const uint32_t nf_ab = (ca >> 2) + choose2((cbb >> 2) + 1);
            const uint32_t sc5 = scb + (1u << (4 * (ca & 3u))) + (1u << (4 * (cbb & 3u)));
            const uint64_t cs5 = bset | card_bit(path_of_code(ca)) | card_bit(path_of_code(cbb));
            const uint32_t wab = S.W[a * kPitch + b];
            const uint32_t *rowA = S.W + a * kPitch, *rowB = S.W + b * kPitch;
            // a flush needs 5 of a suit among 7 cards, i.e. >= 3 among board + a + b (biased nibble >= 6)
            if ((sc5 + 0x2222u) & 0x8888u) run_task<true>(S, lane, n_cd, rowA, rowB, wab, nf_ab, sc5, cs5, acc);
            else run_task<false>(S, lane, n_cd, rowA, rowB, wab, nf_ab, sc5, cs5, acc);
        }
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            wide[0] += acc[q] & 63u; wide[1] += (acc[q] >> 6) & 63u; wide[2] += (acc[q] >> 12) & 63u;
            wide[3] += acc[3 + q] & 63u; wide[4] += (acc[3 + q] >> 6) & 63u; wide[5] += (acc[3 + q] >> 12) & 63u;
        }
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const uint32_t v = __reduce_add_sync(0xFFFFFFFFu, wide[q]);
            if (lane == 0 && v) atomicAdd(&S.cnt[q], v);
        }
        __syncthreads();
        if (tid < 16) {
            uint32_t v;
            if (tid < 3) v = S.cnt[6 + tid];                                  // ahead, tied, behind (HandStrength)
            else if (tid < 12) {
                const uint32_t cls = (tid - 3) / 3, o = (tid - 3) % 3;
                const uint32_t lt = S.cnt[cls], gt = S.cnt[3 + cls], all = 990u * S.cnt[6 + cls];
                v = o == 0 ? lt : (o == 2 ? gt : all - lt - gt);              // HP[class][ahead|tied|behind]
            } else if (tid < 15) v = S.cnt[6 + tid - 12];                     // HPTotal
            else v = 3u;
            out[idx * 16 + tid] = v;
        }
    }
    if (saw_other && tid == 0) atomicOr(other_rows, saw_other);
}

// =====================================================================================================================
// v3: lane = (a,b,d), walking c.  Same counts as the kernel above with ~half the instructions per 4-set:
//   * a static, situation-independent work list (tables.h: flop_task_hdr / flop_lanes) assigns every lane one triple
//     (a,b,d) of unseen positions; the 32 lanes of a warp task share the walk length d-b-1, so there is no ragged tail
//     and no per-iteration activity predicate (padding lanes read all-zero words and therefore add nothing);
//   * everything that depends on (a,b,d) only -- W[a][b], W[a][d], W[b][d], the multiset index prefix, the flush suit
//     of the six fixed cards and its rank mask -- is hoisted out of the walk; per step the lane reads W[a][c], W[b][c],
//     W[c][d] (constant strides), one packed word X[suit][c] = C(rc+2,3) | rank bit of c in that suit << 16, and NF;
//   * the twelve conditional accumulations are issued as predicated IMADs (multiplier 1 from a kernel argument) so
//     that they run on the FMA pipe while the compares run on the ALU pipe.
// =====================================================================================================================
namespace {

constexpr int kPitch3 = 49;        // odd pitch: lanes with consecutive a and equal c hit distinct banks
constexpr int kRows3 = 92;         // 47 real rows + zero rows so that padding lanes can walk 44 rows past row 47
constexpr int kXPitch = 96;

struct Flop3Smem {
    uint16_t flush[kFlushEntries];
    uint32_t W[kRows3 * kPitch3];      // [i][j], i<j<47: our7 << 19 | one-hot class; everything else stays zero
    uint32_t X[4 * kXPitch];           // [suit][i]: C(r_i+2,3) | (suit_i == suit ? 1 << r_i : 0) << 16; i >= 47 zero
    uint32_t I[48];                    // per position: r | C(r+1,2) << 4 | C(r+3,4) << 11 | suit << 22
    uint32_t NF32[kNf + 4];            // non-flush rank << 19 (the compare threshold, ready to use)
    uint16_t Xb[4 * kXPitch];          // [suit][i]: rank bit of i if it has that suit (high half of X, as its own array)
    uint16_t C3x4[kXPitch];            // per position: 4 * C(r_i+2,3) = byte offset contribution into NF32
    uint16_t NF[kNf + 4];
    uint32_t cardkey[64];
    uint32_t cnt[12];
    uint32_t bmask[4];                 // board rank mask per suit
    uint8_t code[64];
};

__device__ __forceinline__ void add_if_lt(uint32_t &acc, uint32_t wq, uint32_t thr, uint32_t wp, uint32_t one) {
    asm("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %1, %2;\n\t@p mad.lo.u32 %0, %3, %4, %0;\n\t}"
        : "+r"(acc) : "r"(wq), "r"(thr), "r"(wp), "r"(one));
}
__device__ __forceinline__ void add_if_ge(uint32_t &acc, uint32_t wq, uint32_t thr, uint32_t wp, uint32_t one) {
    asm("{\n\t.reg .pred p;\n\tsetp.ge.u32 p, %1, %2;\n\t@p mad.lo.u32 %0, %3, %4, %0;\n\t}"
        : "+r"(acc) : "r"(wq), "r"(thr), "r"(wp), "r"(one));
}

}  // namespace

__global__ void __launch_bounds__(kThreads)
hp_flop_treys_v3_kernel(const uint8_t *__restrict__ sit, int64_t n, uint32_t *__restrict__ out, DevTables T,
                        int *__restrict__ other_rows, uint32_t one) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    Flop3Smem &S = *reinterpret_cast<Flop3Smem *>(smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(T.flush);
        uint4 *dst = reinterpret_cast<uint4 *>(S.flush);
        for (uint32_t i = tid; i < kFlushEntries / 8; i += kThreads) dst[i] = src[i];
        for (uint32_t i = tid; i < (uint32_t)(kRows3 * kPitch3); i += kThreads) S.W[i] = 0;   // pad rows stay zero
        for (uint32_t i = tid; i < (uint32_t)(4 * kXPitch); i += kThreads) { S.X[i] = 0; S.Xb[i] = 0; }
        if (tid < (uint32_t)kXPitch) S.C3x4[tid] = 0;
        if (tid < 64) S.cardkey[tid] = tid < 52 ? c_rank_key[tid % 13] : 0u;
    }
    const NfHash h5 = T.nf[0], h7 = T.nf[2];
    const uint32_t n_tasks = T.n_flop_tasks;
    int saw_other = 0;

    for (int64_t idx = blockIdx.x; idx < n; idx += gridDim.x) {
        __syncthreads();
        uint2 w = __ldg(reinterpret_cast<const uint2 *>(sit) + idx);
        uint32_t cb[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) { cb[k] = (w.x >> (8 * k)) & 0xFFu; cb[4 + k] = (w.y >> (8 * k)) & 0xFFu; }
        const bool flop = cb[5] == 0xFFu && cb[6] == 0xFFu;
        if (!flop) { saw_other |= (cb[6] == 0xFFu) ? 1 : 2; continue; }
        bool ok = true;
        uint64_t known = 0, bset = 0;
        uint32_t kb = 0, khole = 0, scb = 0x4444u;   // suit nibbles biased by 4: bit 3 set <=> count >= 4
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const uint32_t c = cb[k];
            ok = ok && c < 52u;
            const uint32_t cc = c < 52u ? c : 0u;
            known |= card_bit(cc);
            if (k >= 2) { bset |= card_bit(cc); kb += S.cardkey[cc]; scb += 1u << (4 * card_suit(cc)); }
            else khole += S.cardkey[cc];
        }
        ok = ok && __popcll(known) == 5;
        if (!ok) {
            if (tid < 16) out[idx * 16 + tid] = 0xFFFFFFFFu;
            continue;
        }
        if (tid < 12) S.cnt[tid] = 0;
        if (tid < 4) S.bmask[tid] = (uint32_t)(bset >> (13 * tid)) & 0x1FFFu;
        if (warp == 0) {   // unseen cards in (rank, suit) order
            const uint32_t p0 = path_of_code(lane), p1 = path_of_code(lane + 32);
            const bool in0 = !((known >> p0) & 1ull);
            const bool in1 = (lane + 32 < 52) && !((known >> p1) & 1ull);
            const uint32_t m0 = __ballot_sync(0xFFFFFFFFu, in0), m1 = __ballot_sync(0xFFFFFFFFu, in1);
            const uint32_t lt = (1u << lane) - 1u;
            if (in0) S.code[__popc(m0 & lt)] = (uint8_t)lane;
            if (in1) S.code[__popc(m0) + __popc(m1 & lt)] = (uint8_t)(lane + 32);
        }
        {   // NF[1820]: board + every multiset of four ranks (rank list from the static table)
            uint32_t bcnt = 0;
#pragma unroll
            for (int k = 2; k < 5; ++k) bcnt += 1u << (2 * card_rank(cb[k]));
            for (uint32_t e = tid; e < (uint32_t)kNf; e += kThreads) {
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
                S.NF32[e] = valid ? nf_lookup(h7, key) << 19 : 0u;
            }
        }
        __syncthreads();
        {   // per-position tables
            if (tid < (uint32_t)kUnseen) {
                const uint32_t cd = S.code[tid], r = cd >> 2, su = cd & 3u;
                S.I[tid] = r | (choose2(r + 1) << 4) | (choose4(r + 3) << 11) | (su << 22);
#pragma unroll
                for (uint32_t s2 = 0; s2 < 4; ++s2) {
                    S.X[s2 * kXPitch + tid] = choose3(r + 2) | ((su == s2 ? (1u << r) : 0u) << 16);
                    S.Xb[s2 * kXPitch + tid] = (uint16_t)(su == s2 ? (1u << r) : 0u);
                }
                S.C3x4[tid] = (uint16_t)(4u * choose3(r + 2));
            }
            const uint32_t ours_now = treys_rank(h5, S.flush, kb + khole, known);
            uint32_t c0 = 0, c1 = 0, c2 = 0;
            for (uint32_t p = tid; p < (uint32_t)kPairs; p += kThreads) {
                const uint32_t pr = __ldg(T.pairs + p);
                const uint32_t i = pr & 0xFFu, j = pr >> 8;
                const uint32_t pi = path_of_code(S.code[i]), pj = path_of_code(S.code[j]);
                const uint32_t key = S.cardkey[pi] + S.cardkey[pj];
                const uint64_t cs = card_bit(pi) | card_bit(pj);
                const uint32_t our7 = treys_rank(h7, S.flush, kb + khole + key, known | cs);
                const uint32_t opp5 = treys_rank(h5, S.flush, kb + key, bset | cs);
                const uint32_t cls = ours_now < opp5 ? 0u : (ours_now == opp5 ? 1u : 2u);
                c0 += cls == 0; c1 += cls == 1; c2 += cls == 2;
                S.W[i * kPitch3 + j] = (our7 << 19) | (1u << (6 * cls));
            }
            c0 = __reduce_add_sync(0xFFFFFFFFu, c0);
            c1 = __reduce_add_sync(0xFFFFFFFFu, c1);
            c2 = __reduce_add_sync(0xFFFFFFFFu, c2);
            if (lane == 0) { atomicAdd(&S.cnt[6], c0); atomicAdd(&S.cnt[7], c1); atomicAdd(&S.cnt[8], c2); }
        }
        __syncthreads();

        uint32_t acc[6] = {0, 0, 0, 0, 0, 0};
        uint32_t wide[6] = {0, 0, 0, 0, 0, 0};
        uint32_t since = 0;
        uint32_t next_lane = warp < n_tasks ? __ldg(T.flop_lanes + warp * 32 + lane) : 0xFFFFFFFFu;
        uint32_t next_hdr = warp < n_tasks ? __ldg(T.flop_task_hdr + warp) : 0u;
        for (uint32_t t = warp; t < n_tasks; t += kWarps) {
            const uint32_t trip = next_lane, hdr = next_hdr;
            if (t + kWarps < n_tasks) {   // prefetch the next task of this warp
                next_lane = __ldg(T.flop_lanes + (t + kWarps) * 32 + lane);
                next_hdr = __ldg(T.flop_task_hdr + t + kWarps);
            }
            const uint32_t k0 = hdr & 0xFFu, iters = hdr >> 8;
            if (iters == 0) continue;
            if (since + iters > 31) {
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    wide[0] += acc[q] & 63u; wide[1] += (acc[q] >> 6) & 63u; wide[2] += (acc[q] >> 12) & 63u;
                    wide[3] += acc[3 + q] & 63u; wide[4] += (acc[3 + q] >> 6) & 63u; wide[5] += (acc[3 + q] >> 12) & 63u;
                    acc[q] = 0; acc[3 + q] = 0;
                }
                since = 0;
            }
            since += iters;
            // ---- lane invariants ---------------------------------------------------------------------------------------
            uint32_t wab = 0, wad = 0, wbd = 0, nf = 0, mask6 = 0;
            const uint32_t *pac = S.W + 47 * kPitch3, *pbc = pac, *pcd = pac;   // zero words
            const uint16_t *pxb = S.Xb + 48, *pc3 = S.C3x4 + 48;
            if (trip != 0xFFFFFFFFu) {
                const uint32_t a = trip & 0xFFu, b = (trip >> 8) & 0xFFu, d = (trip >> 16) & 0xFFu;
                const uint32_t ia = S.I[a], ib = S.I[b], id = S.I[d];
                nf = (ia & 15u) + ((ib >> 4) & 127u) + ((id >> 11) & 2047u);
                const uint32_t sa = ia >> 22, sb = ib >> 22, sd = id >> 22;
                const uint32_t sc6 = scb + (1u << (4 * sa)) + (1u << (4 * sb)) + (1u << (4 * sd));
                const uint32_t fl = sc6 & 0x8888u;
                uint32_t sf = 0;
                if (fl) {
                    sf = (uint32_t)(__ffs(fl) - 1) >> 2;
                    const uint32_t *xr = S.X + sf * kXPitch;
                    mask6 = S.bmask[sf] | ((xr[a] | xr[b] | xr[d]) >> 16);
                }
                wab = S.W[a * kPitch3 + b];
                wad = S.W[a * kPitch3 + d];
                wbd = S.W[b * kPitch3 + d];
                const uint32_t c0i = b + 1 + k0;
                pac = S.W + a * kPitch3 + c0i;
                pbc = S.W + b * kPitch3 + c0i;
                pcd = S.W + c0i * kPitch3 + d;
                pxb = S.Xb + sf * kXPitch + c0i;
                pc3 = S.C3x4 + c0i;
                nf *= 4u;   // byte offset into NF32
            }
            // ---- the walk over c --------------------------------------------------------------------------------------------
#pragma unroll 4
            for (uint32_t k = 0; k < iters; ++k) {
                const uint32_t wac = *pac, wbc = *pbc, wcd = *pcd;
                const uint32_t xb = *pxb, c3 = *pc3;
                ++pac; ++pbc; ++pxb; ++pc3; pcd += kPitch3;
                uint32_t lo = *reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(S.NF32) + nf + c3);
                const uint32_t m = mask6 | xb;
                if (__popc(m) >= 5) lo = (uint32_t)S.flush[m] << 19;
                const uint32_t hi = lo + (1u << 19);
                add_if_lt(acc[0], wcd, lo, wab, one);  add_if_ge(acc[3], wcd, hi, wab, one);   // P=ab Q=cd
                add_if_lt(acc[0], wab, lo, wcd, one);  add_if_ge(acc[3], wab, hi, wcd, one);   // P=cd Q=ab
                add_if_lt(acc[1], wbd, lo, wac, one);  add_if_ge(acc[4], wbd, hi, wac, one);   // P=ac Q=bd
                add_if_lt(acc[1], wac, lo, wbd, one);  add_if_ge(acc[4], wac, hi, wbd, one);   // P=bd Q=ac
                add_if_lt(acc[2], wbc, lo, wad, one);  add_if_ge(acc[5], wbc, hi, wad, one);   // P=ad Q=bc
                add_if_lt(acc[2], wad, lo, wbc, one);  add_if_ge(acc[5], wad, hi, wbc, one);   // P=bc Q=ad
            }
        }
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            wide[0] += acc[q] & 63u; wide[1] += (acc[q] >> 6) & 63u; wide[2] += (acc[q] >> 12) & 63u;
            wide[3] += acc[3 + q] & 63u; wide[4] += (acc[3 + q] >> 6) & 63u; wide[5] += (acc[3 + q] >> 12) & 63u;
        }
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const uint32_t v = __reduce_add_sync(0xFFFFFFFFu, wide[q]);
            if (lane == 0 && v) atomicAdd(&S.cnt[q], v);
        }
        __syncthreads();
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

// Second pass for rows the flop kernel skipped: the plain-enumeration kernel filtered to non-flop rows.
cudaError_t launch_hp_flop_treys(const DevTables &T, const uint8_t *sit, int64_t n, uint32_t *out, int sms,
                                 int *other_rows_flag, int impl, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    if (impl != 2) {
        const size_t smem3 = sizeof(Flop3Smem);
        cudaError_t e3 = cudaFuncSetAttribute(hp_flop_treys_v3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3);
        if (e3 != cudaSuccess) return e3;
        static int bps3 = 0;
        if (bps3 == 0) {
            e3 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps3, hp_flop_treys_v3_kernel, kThreads, smem3);
            if (e3 != cudaSuccess) return e3;
            if (bps3 < 1) bps3 = 1;
        }
        int64_t blocks3 = (int64_t)sms * bps3;
        if (blocks3 > n) blocks3 = n;
        e3 = cudaMemsetAsync(other_rows_flag, 0, sizeof(int), st);
        if (e3 != cudaSuccess) return e3;
        count_launch(); hp_flop_treys_v3_kernel<<<(int)blocks3, kThreads, smem3, st>>>(sit, n, out, T, other_rows_flag, 1u);
        return cudaGetLastError();
    }
    const size_t smem = sizeof(FlopSmem);
    cudaError_t e = cudaFuncSetAttribute(hp_flop_treys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    static int blocks_per_sm = 0;
    if (blocks_per_sm == 0) {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, hp_flop_treys_kernel, kThreads, smem);
        if (e != cudaSuccess) return e;
        if (blocks_per_sm < 1) blocks_per_sm = 1;
    }
    int64_t blocks = (int64_t)sms * blocks_per_sm;
    if (blocks > n) blocks = n;
    e = cudaMemsetAsync(other_rows_flag, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    count_launch(); hp_flop_treys_kernel<<<(int)blocks, kThreads, smem, st>>>(sit, n, out, T, other_rows_flag);
    return cudaGetLastError();
}

}  // namespace holdem
