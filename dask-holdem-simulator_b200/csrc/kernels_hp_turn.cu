// kernels_hp_turn.cu -- turn hand strength + hand potential in Treys mode (HandCalculations.py:178-226 with one river card
// to come), one CTA per situation.  sm_100a only.
//
// Same idea as the flop kernel (kernels_hp_flop.cu) one dimension down.  The reference-shaped loop visits 1035 opponent
// holdings x 44 rivers = 45,540 pairs; the opponent's final hand is board4 + P + r and depends only on the 3-SET
// {o1,o2,r} of unseen cards: C(46,3) = 15,180 sets, each reached by 3 splits.  Per situation the kernel evaluates
// 15,180 opponent hands + 46 own 7-card hands + 1,036 six-card hands + 455 table entries and applies
//        HP[class(P)][cmp(our7[r], R(S))] += 1      for the 3 ways to write S = P + r.
// Lane = an opponent pair position (a,b) from the static work list (tables.h: turn_lanes); it walks c = b+1..45.
#include "device_common.cuh"
#include "launch.h"

namespace holdem {

namespace {

constexpr int kTThreads = 128;
constexpr int kTWarps = kTThreads / 32;
constexpr int kTUnseen = 46;
constexpr int kTPairs = 1035;      // C(46,2)
constexpr int kTPitch = 47;        // odd pitch
constexpr int kTRows = 92;         // real rows + zero rows for padding lanes
constexpr int kTNf = 455;          // multisets of 3 ranks

struct TurnSmem {
    uint16_t flush[kFlushEntries];
    uint32_t W[kTRows * kTPitch];      // [i][j], i<j<46: one-hot class word of the pair (1 << 6*class); else zero
    uint32_t O[96];                    // per position: our7 << 19 with that river card; entries >= 46 zero
    uint32_t X[4 * 96];                // [suit][i]: C(r_i+2,3) | rank bit of i if it has that suit << 16
    uint32_t I[48];                    // per position: r | C(r+1,2) << 4 | suit << 22
    uint16_t NF[kTNf + 1];
    uint32_t cardkey[64];
    uint32_t cnt[12];
    uint32_t bmask[4];
    uint8_t code[64];
};

__device__ __forceinline__ uint32_t t_path_of_code(uint32_t code) { return (code & 3u) * 13u + (code >> 2); }
__device__ __forceinline__ uint32_t t_choose2(uint32_t n) { return n * (n - 1) / 2; }
__device__ __forceinline__ uint32_t t_choose3(uint32_t n) { return n * (n - 1) * (n - 2) / 6; }

__device__ __forceinline__ void t_add_if_lt(uint32_t &acc, uint32_t wq, uint32_t thr, uint32_t wp, uint32_t one) {
    asm("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %1, %2;\n\t@p mad.lo.u32 %0, %3, %4, %0;\n\t}"
        : "+r"(acc) : "r"(wq), "r"(thr), "r"(wp), "r"(one));
}
__device__ __forceinline__ void t_add_if_ge(uint32_t &acc, uint32_t wq, uint32_t thr, uint32_t wp, uint32_t one) {
    asm("{\n\t.reg .pred p;\n\tsetp.ge.u32 p, %1, %2;\n\t@p mad.lo.u32 %0, %3, %4, %0;\n\t}"
        : "+r"(acc) : "r"(wq), "r"(thr), "r"(wp), "r"(one));
}

}  // namespace

__global__ void __launch_bounds__(kTThreads)
hp_turn_treys_kernel(const uint8_t *__restrict__ sit, int64_t n, uint32_t *__restrict__ out, DevTables T,
                     const int *__restrict__ run_flag, uint32_t one) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    TurnSmem &S = *reinterpret_cast<TurnSmem *>(smem_raw);
    if (run_flag != nullptr && (*run_flag & 1) == 0) return;   // the flop kernel saw no 4-card boards
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(T.flush);
        uint4 *dst = reinterpret_cast<uint4 *>(S.flush);
        for (uint32_t i = tid; i < kFlushEntries / 8; i += kTThreads) dst[i] = src[i];
        for (uint32_t i = tid; i < (uint32_t)(kTRows * kTPitch); i += kTThreads) S.W[i] = 0;
        for (uint32_t i = tid; i < 4 * 96; i += kTThreads) S.X[i] = 0;
        if (tid < 96) S.O[tid] = 0;
        if (tid < 64) S.cardkey[tid] = tid < 52 ? c_rank_key[tid % 13] : 0u;
    }
    const NfHash h6 = T.nf[1], h7 = T.nf[2];
    const uint32_t n_tasks = T.n_turn_tasks;

    for (int64_t idx = blockIdx.x; idx < n; idx += gridDim.x) {
        __syncthreads();
        uint2 w = __ldg(reinterpret_cast<const uint2 *>(sit) + idx);
        uint32_t cb[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) { cb[k] = (w.x >> (8 * k)) & 0xFFu; cb[4 + k] = (w.y >> (8 * k)) & 0xFFu; }
        if (!(cb[5] != 0xFFu && cb[6] == 0xFFu)) continue;   // not a turn row
        bool ok = true;
        uint64_t known = 0, bset = 0;
        uint32_t kb = 0, khole = 0, scb = 0x4444u;   // suit nibbles biased by 4: bit 3 set <=> count >= 4
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const uint32_t c = cb[k];
            ok = ok && c < 52u;
            const uint32_t cc = c < 52u ? c : 0u;
            known |= card_bit(cc);
            if (k >= 2) { bset |= card_bit(cc); kb += S.cardkey[cc]; scb += 1u << (4 * card_suit(cc)); }
            else khole += S.cardkey[cc];
        }
        ok = ok && __popcll(known) == 6;
        if (!ok) {
            if (tid < 16) out[idx * 16 + tid] = 0xFFFFFFFFu;
            continue;
        }
        if (tid < 12) S.cnt[tid] = 0;
        if (tid < 4) S.bmask[tid] = (uint32_t)(bset >> (13 * tid)) & 0x1FFFu;
        if (warp == 0) {   // unseen cards in (rank, suit) order
            const uint32_t p0 = t_path_of_code(lane), p1 = t_path_of_code(lane + 32);
            const bool in0 = !((known >> p0) & 1ull);
            const bool in1 = (lane + 32 < 52) && !((known >> p1) & 1ull);
            const uint32_t m0 = __ballot_sync(0xFFFFFFFFu, in0), m1 = __ballot_sync(0xFFFFFFFFu, in1);
            const uint32_t lt = (1u << lane) - 1u;
            if (in0) S.code[__popc(m0 & lt)] = (uint8_t)lane;
            if (in1) S.code[__popc(m0) + __popc(m1 & lt)] = (uint8_t)(lane + 32);
        }
        {   // NF[455]: board4 + every multiset of three ranks
            uint64_t bcnt = 0;   // 13 x 4-bit board rank multiplicities (a turn board can hold 4 of a rank)
#pragma unroll
            for (int k = 2; k < 6; ++k) bcnt += 1ull << (4 * card_rank(cb[k]));
            for (uint32_t e = tid; e < (uint32_t)kTNf; e += kTThreads) {
                const uint32_t rr = __ldg(T.nf3_ranks + e);
                uint32_t key = kb;
                bool valid = true;
                uint32_t prev = 99, run = 0;
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const uint32_t r = (rr >> (4 * q)) & 15u;
                    key += S.cardkey[r];
                    run = (r == prev) ? run + 1 : 1;
                    prev = r;
                    const uint32_t onboard = (uint32_t)(bcnt >> (4 * r)) & 15u;
                    valid = valid && (run + onboard <= 4);
                }
                S.NF[e] = valid ? (uint16_t)nf_lookup(h7, key) : (uint16_t)0;
            }
        }
        __syncthreads();
        {   // per-position and per-pair tables
            if (tid < (uint32_t)kTUnseen) {
                const uint32_t cd = S.code[tid], r = cd >> 2, su = cd & 3u;
                const uint32_t p = t_path_of_code(cd);
                S.I[tid] = r | (t_choose2(r + 1) << 4) | (su << 22);
#pragma unroll
                for (uint32_t s2 = 0; s2 < 4; ++s2)
                    S.X[s2 * 96 + tid] = t_choose3(r + 2) | ((su == s2 ? (1u << r) : 0u) << 16);
                S.O[tid] = treys_rank(h7, S.flush, kb + khole + S.cardkey[p], known | card_bit(p)) << 19;
            }
            const uint32_t ours_now = treys_rank(h6, S.flush, kb + khole, known);
            uint32_t c0 = 0, c1 = 0, c2 = 0;
            for (uint32_t p = tid; p < (uint32_t)kTPairs; p += kTThreads) {
                const uint32_t pr = __ldg(T.pairs + p);
                const uint32_t i = pr & 0xFFu, j = pr >> 8;
                const uint32_t pi = t_path_of_code(S.code[i]), pj = t_path_of_code(S.code[j]);
                const uint32_t opp6 = treys_rank(h6, S.flush, kb + S.cardkey[pi] + S.cardkey[pj],
                                                 bset | card_bit(pi) | card_bit(pj));
                const uint32_t cls = ours_now < opp6 ? 0u : (ours_now == opp6 ? 1u : 2u);
                c0 += cls == 0; c1 += cls == 1; c2 += cls == 2;
                S.W[i * kTPitch + j] = 1u << (6 * cls);
            }
            c0 = __reduce_add_sync(0xFFFFFFFFu, c0);
            c1 = __reduce_add_sync(0xFFFFFFFFu, c1);
            c2 = __reduce_add_sync(0xFFFFFFFFu, c2);
            if (lane == 0) { atomicAdd(&S.cnt[6], c0); atomicAdd(&S.cnt[7], c1); atomicAdd(&S.cnt[8], c2); }
        }
        __syncthreads();

        uint32_t acc[6] = {0, 0, 0, 0, 0, 0};      // packed 3 x 6-bit class counters: lt x3 splits, gt x3 splits
        uint32_t wide[6] = {0, 0, 0, 0, 0, 0};
        uint32_t since = 0;
        for (uint32_t t = warp; t < n_tasks; t += kTWarps) {
            const uint32_t hdr = __ldg(T.turn_task_hdr + t);
            const uint32_t trip = __ldg(T.turn_lanes + t * 32 + lane);
            const uint32_t k0 = hdr & 0xFFu, iters = hdr >> 8;
            if (iters == 0) continue;
            if (since + iters > 63) {   // one add per step and field: drain before 63
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    wide[0] += acc[q] & 63u; wide[1] += (acc[q] >> 6) & 63u; wide[2] += (acc[q] >> 12) & 63u;
                    wide[3] += acc[3 + q] & 63u; wide[4] += (acc[3 + q] >> 6) & 63u; wide[5] += (acc[3 + q] >> 12) & 63u;
                    acc[q] = 0; acc[3 + q] = 0;
                }
                since = 0;
            }
            since += iters;
            uint32_t wab = 0, oa = 0, ob = 0, nf = 0, mask6 = 0;
            const uint32_t *pac = S.W + 46 * kTPitch, *pbc = pac, *px = S.X + 48, *po = S.O + 48;   // zero words
            if (trip != 0xFFFFFFFFu) {
                const uint32_t a = trip & 0xFFu, b = (trip >> 8) & 0xFFu;
                const uint32_t ia = S.I[a], ib = S.I[b];
                nf = (ia & 15u) + ((ib >> 4) & 127u);
                const uint32_t sc6 = scb + (1u << (4 * (ia >> 22))) + (1u << (4 * (ib >> 22)));
                const uint32_t fl = sc6 & 0x8888u;
                uint32_t sf = 0;
                if (fl) {
                    sf = (uint32_t)(__ffs(fl) - 1) >> 2;
                    const uint32_t *xr = S.X + sf * 96;
                    mask6 = S.bmask[sf] | ((xr[a] | xr[b]) >> 16);
                }
                wab = S.W[a * kTPitch + b];
                oa = S.O[a];
                ob = S.O[b];
                const uint32_t c0i = b + 1 + k0;
                pac = S.W + a * kTPitch + c0i;
                pbc = S.W + b * kTPitch + c0i;
                px = S.X + sf * 96 + c0i;
                po = S.O + c0i;
            }
#pragma unroll 2
            for (uint32_t k = 0; k < iters; ++k) {
                const uint32_t wac = pac[k], wbc = pbc[k], x = px[k], oc = po[k];
                uint32_t R = S.NF[nf + (x & 0xFFFFu)];
                const uint32_t m = mask6 | (x >> 16);
                if (__popc(m) >= 5) R = S.flush[m];
                const uint32_t lo = R << 19, hi = lo + (1u << 19);
                t_add_if_lt(acc[0], oc, lo, wab, one);  t_add_if_ge(acc[3], oc, hi, wab, one);   // P=ab river=c
                t_add_if_lt(acc[1], ob, lo, wac, one);  t_add_if_ge(acc[4], ob, hi, wac, one);   // P=ac river=b
                t_add_if_lt(acc[2], oa, lo, wbc, one);  t_add_if_ge(acc[5], oa, hi, wbc, one);   // P=bc river=a
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
                const uint32_t lt = S.cnt[cls], gt = S.cnt[3 + cls], all = 44u * S.cnt[6 + cls];
                v = o == 0 ? lt : (o == 2 ? gt : all - lt - gt);
            } else if (tid < 15) v = S.cnt[6 + tid - 12];
            else v = 4u;
            out[idx * 16 + tid] = v;
        }
    }
}

cudaError_t launch_hp_turn_treys(const DevTables &T, const uint8_t *sit, int64_t n, uint32_t *out, int sms,
                                 const int *run_flag, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const size_t smem = sizeof(TurnSmem);
    cudaError_t e = cudaFuncSetAttribute(hp_turn_treys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    static int bps = 0;
    if (bps == 0) {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, hp_turn_treys_kernel, kTThreads, smem);
        if (e != cudaSuccess) return e;
        if (bps < 1) bps = 1;
    }
    int64_t blocks = (int64_t)sms * bps;
    if (blocks > n) blocks = n;
    count_launch(); hp_turn_treys_kernel<<<(int)blocks, kTThreads, smem, st>>>(sit, n, out, T, run_flag, 1u);
    return cudaGetLastError();
}

}  // namespace holdem
