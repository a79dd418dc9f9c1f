// kernels_eval.cu -- batched 5/6/7-card Treys evaluator, literal category ranker, exhaustive 7-card sweep and the
// shared-memory gather micro-benchmark.  sm_100a only.
//
// Replaces treys.Evaluator.evaluate (reached from HandCalculations.py:229-240, pokerlib_to_treys.py:15-21) and
// CalculateHandRankValue (HandCalculations.py:45-82) for batches of hands.
#include <cstdlib>

#include "device_common.cuh"
#include "launch.h"

namespace holdem {

// =====================================================================================================================
// Batched evaluator.  One hand per thread per grid-stride step; the input is one coalesced 8-byte load per hand, the
// output one 2-byte store: 10 algorithmic HBM bytes per evaluation.
// kShared: the flush table, the displacement table and the value table of the hand size are staged in shared memory
// (16 KB + 4..32 KB + 16..128 KB) once per persistent CTA; otherwise they are read through L1/L2 (small batches).
// =====================================================================================================================
template <int NC>
__device__ __forceinline__ uint32_t eval_row(uint2 w, const uint32_t *cardkey, const uint16_t *flush,
                                             const NfHash &h) {
    uint32_t key = 0, lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        uint32_t c = (k < 4) ? ((w.x >> (8 * k)) & 0xFFu) : ((w.y >> (8 * (k - 4))) & 0xFFu);
        key += cardkey[c & 63u];
        lo |= shl_clamp(1u, c);
        hi |= shl_clamp(1u, c - 32u);
    }
    // duplicates or cards > 51 leave fewer than NC bits below bit 52
    bool bad = (__popc(lo) + __popc(hi & 0xFFFFFu) != NC) || (hi >> 20);
    uint64_t cards = ((uint64_t)hi << 32) | lo;
    uint32_t r = treys_rank(h, flush, key, cards);
    return bad ? 0u : r;
}

// Strict path (and the small-batch path): builds the 52-bit card set, so duplicate cards are detected.
template <int NC, bool kShared>
__global__ void __launch_bounds__(kShared ? 1024 : 256, 1)
eval_batch_kernel(const uint2 *__restrict__ cards, int64_t n, uint16_t *__restrict__ out, DevTables T) {
    extern __shared__ __align__(16) uint8_t smem[];
    const NfHash &g = T.nf[NC - 5];
    NfHash h = g;
    const uint16_t *flush = T.flush;
    uint32_t *s_cardkey = reinterpret_cast<uint32_t *>(smem);                 // 64 entries
    if (threadIdx.x < 64) s_cardkey[threadIdx.x] = threadIdx.x < 52 ? c_rank_key[threadIdx.x % 13] : 0u;
    if (kShared) {
        uint16_t *s_flush = reinterpret_cast<uint16_t *>(smem + 1024);
        uint16_t *s_disp = s_flush + kFlushEntries;
        uint16_t *s_val = s_disp + T.disp_entries[NC - 5];
        // 16-byte vector copies; every section is 16-byte aligned and a multiple of 16 bytes long
        const uint4 *src;
        uint4 *dst;
        src = reinterpret_cast<const uint4 *>(T.flush); dst = reinterpret_cast<uint4 *>(s_flush);
        for (uint32_t i = threadIdx.x; i < kFlushEntries / 8; i += blockDim.x) dst[i] = src[i];
        src = reinterpret_cast<const uint4 *>(g.disp); dst = reinterpret_cast<uint4 *>(s_disp);
        for (uint32_t i = threadIdx.x; i < T.disp_entries[NC - 5] / 8; i += blockDim.x) dst[i] = src[i];
        src = reinterpret_cast<const uint4 *>(g.val); dst = reinterpret_cast<uint4 *>(s_val);
        for (uint32_t i = threadIdx.x; i < T.val_entries[NC - 5] / 8; i += blockDim.x) dst[i] = src[i];
        h.disp = s_disp;
        h.val = s_val;
        flush = s_flush;
    }
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint2 w = __ldg(cards + i);
        out[i] = (uint16_t)eval_row<NC>(w, s_cardkey, flush, h);
    }
}

// Fast path for large batches.  The per-card table word carries the additive rank key in its low 23 bits and three
// 3-bit suit counters above it (clubs, diamonds, hearts; spades = NC - the rest), so one shared-memory gather and one
// add per card yield both the rank multiset key and the suit histogram; no card set is built.  Hands that hold a flush
// (about 3 % of 7-card hands) or an out-of-range byte are deferred and finished by eval_row() 32 at a time.  Duplicate
// cards are NOT detected here (neither does Treys); the strict kernel above does that.
//  * the card table is [52 rows of 256 bytes][32 lane words]: one PRMT builds lane * 4 | byte << 8, the whole byte offset,
//    so a card costs a permute, a gather and a multiply-add (no address multiply);
//  * a deferred hand (flush pattern or out-of-range byte) only sets a bit in a per-lane 32-bit window (bit = hand number
//    inside the window of eight passes): one predicated multiply-add, no vote, no ballot;
//  * every eight passes the warp compacts its windows into a shared-memory stack of hand numbers (prefix sum over the
//    lanes) and evaluates them 32 at a time with eval_row(); what is left (< 32) stays on the stack for the next window.
constexpr uint32_t kTab2Bytes = 52 * 256;
constexpr int kStackCap = 160;   // hand numbers per warp: up to 31 carried over + one window's worth that still fits

template <int NC>
__global__ void __launch_bounds__(1024, 1)
eval_batch_fast_kernel(const uint2 *__restrict__ cards, int64_t n, uint16_t *__restrict__ out, DevTables T,
                        uint32_t one) {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int kIlp = 4;
    constexpr uint32_t kSmask = NC == 7 ? 0xFFFFu : (NC == 6 ? 0x7FFFu : 0x1FFFu);   // slot bits 16 / 15 / 13 (tables.cpp)
    uint32_t *s_cardkey = reinterpret_cast<uint32_t *>(smem);                 // 64 entries (strict path)
    // Bytes above 51 index past the table -- at most 256 + 255 * 256 + 124 bytes into this CTA's shared memory, which the
    // launcher keeps inside the allocation -- and the hand goes to the strict path anyway.
    uint8_t *s_tab = smem + 256;
    uint32_t *s_slow = reinterpret_cast<uint32_t *>(s_tab + kTab2Bytes);      // 512 bits: suit-counter pattern holds a flush
    uint32_t *s_stack = s_slow + 16;                                          // [32 warps][kStackCap]
    uint16_t *s_flush = reinterpret_cast<uint16_t *>(s_stack + 32 * kStackCap);
    uint16_t *s_disp = s_flush + kFlushEntries;
    uint16_t *s_val = s_disp + T.disp_entries[NC - 5];
    const NfHash &g = T.nf[NC - 5];
    if (threadIdx.x < 64) s_cardkey[threadIdx.x] = threadIdx.x < 52 ? c_rank_key[threadIdx.x % 13] : 0u;
    for (uint32_t i = threadIdx.x; i < 52 * 32; i += blockDim.x) {
        const uint32_t c = i >> 5, su = c / 13;
        *reinterpret_cast<uint32_t *>(s_tab + c * 256 + (i & 31u) * 4) =
            c_rank_key[c % 13] | (su < 3 ? 1u << (23 + 3 * su) : 0u);
    }
    if (threadIdx.x < 16) {
        uint32_t word = 0;
        for (uint32_t b = 0; b < 32; ++b) {
            const uint32_t sv = threadIdx.x * 32 + b, c0 = sv & 7u, c1 = (sv >> 3) & 7u, c2 = sv >> 6;
            if (c0 >= 5 || c1 >= 5 || c2 >= 5 || c0 + c1 + c2 <= (uint32_t)(NC - 5)) word |= 1u << b;
        }
        s_slow[threadIdx.x] = word;
    }
    {
        const uint4 *src;
        uint4 *dst;
        src = reinterpret_cast<const uint4 *>(T.flush); dst = reinterpret_cast<uint4 *>(s_flush);
        for (uint32_t i = threadIdx.x; i < kFlushEntries / 8; i += blockDim.x) dst[i] = src[i];
        src = reinterpret_cast<const uint4 *>(g.disp); dst = reinterpret_cast<uint4 *>(s_disp);
        for (uint32_t i = threadIdx.x; i < T.disp_entries[NC - 5] / 8; i += blockDim.x) dst[i] = src[i];
        src = reinterpret_cast<const uint4 *>(g.val); dst = reinterpret_cast<uint4 *>(s_val);
        for (uint32_t i = threadIdx.x; i < T.val_entries[NC - 5] / 8; i += blockDim.x) dst[i] = src[i];
    }
    NfHash h = g;
    h.disp = s_disp;
    h.val = s_val;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, lane4 = lane * 4;
    uint32_t *stack = s_stack + warp * kStackCap;
    // 32-bit hand indices: the launcher splits batches of 2^31 hands or more
    const uint32_t nn = (uint32_t)n, stride = gridDim.x * blockDim.x;
    // lanes past the end of the batch reload the last hand (clamped index) and neither store nor defer it
    const uint32_t last = nn - 1;

    uint32_t sn = 0;              // entries on this warp's stack (warp-uniform)
    uint32_t window = 0;          // bit k: hand wbase + k * stride of this lane is deferred
    uint32_t itbit = 1;           // 1 << 4 * (pass inside the window)
    uint32_t wbase = blockIdx.x * blockDim.x + threadIdx.x;
    auto slow_eval = [&](uint32_t i) {
        if (i < nn) out[i] = (uint16_t)eval_row<NC>(__ldg(cards + i), s_cardkey, s_flush, h);
    };
    auto close_window = [&](bool last) {
        // exclusive prefix sum of the per-lane counts
        const uint32_t cnt = __popc(window);
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if ((int)lane >= d) incl += t;
        }
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        if (total) {
            if (sn + total > (uint32_t)kStackCap) {
                // a window this full (> 4 deferred hands per lane on average) is not worth compacting
                for (uint32_t bits = window; bits; bits &= bits - 1) slow_eval(wbase + (__ffs(bits) - 1) * stride);
            } else {
                uint32_t pos = sn + incl - cnt;
                for (uint32_t bits = window; bits; bits &= bits - 1) stack[pos++] = wbase + (__ffs(bits) - 1) * stride;
                sn += total;
                __syncwarp();
                while (sn >= 32) {
                    sn -= 32;
                    slow_eval(stack[sn + lane]);
                }
            }
        }
        if (last) {
            __syncwarp();
            if (lane < sn) slow_eval(stack[lane]);
        }
        __syncwarp();
    };

    uint2 wnext[kIlp];
#pragma unroll
    for (int j = 0; j < kIlp; ++j) {
        wnext[j] = __ldg(cards + min(wbase + j * stride, last));
    }
    for (uint32_t base = wbase; base - threadIdx.x < nn; base += kIlp * stride) {
        // the loop bound is block-uniform; the words of this pass were requested one pass ago, the next pass's are requested now
        uint2 w[kIlp];
#pragma unroll
        for (int j = 0; j < kIlp; ++j) {
            w[j] = wnext[j];
            wnext[j] = __ldg(cards + min(base + (kIlp + j) * stride, last));
        }
#pragma unroll
        for (int j = 0; j < kIlp; ++j) {
            const uint32_t i = base + j * stride;
            uint32_t sum = 0;
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                // result bytes: [0] = lane * 4, [1] = card byte k, [2] = [3] = 0
                const uint32_t off = __byte_perm(k < 4 ? w[j].x : w[j].y, lane4, 0x5504u | ((k & 3) << 4));
                const uint32_t v = *reinterpret_cast<const uint32_t *>(s_tab + off);
                // multiply-add by a kernel argument equal to 1: issues on the FMA pipe, not the ALU pipe
                asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(sum) : "r"(v), "r"(one));
            }
            // any byte >= 52?  (b & 0x7F) + 76 sets bit 7 iff (b & 0x7F) >= 52; OR b itself catches b >= 128
            const uint32_t ymask = NC == 7 ? 0x00808080u : (NC == 6 ? 0x00008080u : 0x00000080u);
            const uint32_t badx = (((w[j].x & 0x7F7F7F7Fu) + 0x4C4C4C4Cu) | w[j].x) & 0x80808080u;
            const uint32_t bady = (((w[j].y & 0x7F7F7F7Fu) + 0x4C4C4C4Cu) | w[j].y) & ymask;
            const uint32_t s = sum >> 23;
            const uint32_t flushbit = (s_slow[s >> 5] >> (s & 31u)) & 1u;
            const bool slow = (flushbit | badx | bady) != 0;
            const uint32_t key = sum & 0x7FFFFFu;
            const uint32_t b = (key * h.mul_b) >> h.bshift;
            const uint32_t slot = (((key * h.mul_s) >> h.sshift) + s_disp[b]) & kSmask;
            const uint32_t r = s_val[slot];
            if (i < nn && !slow) out[i] = (uint16_t)r;
            if (slow) window += itbit << j;
        }
        itbit <<= 4;
        if (itbit == 0) {
            close_window(false);
            window = 0;
            itbit = 1;
            wbase = base + kIlp * stride;
        }
    }
    close_window(true);
}

size_t eval_smem_bytes(const DevTables &T, int nc) {
    return 1024 + 2 * (size_t)(kFlushEntries + T.disp_entries[nc - 5] + T.val_entries[nc - 5]);
}
size_t eval_fast_smem_bytes(const DevTables &T, int nc) {
    size_t b = 256 + kTab2Bytes + 64 + 32 * kStackCap * 4 + 2 * (size_t)(kFlushEntries + T.disp_entries[nc - 5] + T.val_entries[nc - 5]);
    const size_t reach = 256 + 256 * 256;   // where a card byte of 255 lands
    return b < reach ? reach : b;
}

template <int NC>
static cudaError_t launch_eval_nc(const DevTables &T, const uint8_t *cards, int64_t n, uint16_t *out, int sms,
                                  bool strict, cudaStream_t st) {
    const uint2 *c2 = reinterpret_cast<const uint2 *>(cards);
    if (n >= (int64_t)sms * 2048) {  // enough work to amortise staging ~50-180 KB per CTA
        if (strict) {
            size_t smem = eval_smem_bytes(T, NC);
            cudaError_t e = cudaFuncSetAttribute(eval_batch_kernel<NC, true>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            count_launch(); eval_batch_kernel<NC, true><<<sms, 1024, smem, st>>>(c2, n, out, T);
        } else {
            // the kernel masks value-table slots with a compile-time constant (16 / 15 / 13 slot bits, tables.cpp)
            if (T.val_entries[NC - 5] != (NC == 7 ? 65536u : (NC == 6 ? 32768u : 8192u))) return cudaErrorInvalidValue;
            const size_t smem = eval_fast_smem_bytes(T, NC);
            cudaError_t e = cudaFuncSetAttribute(eval_batch_fast_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)smem);
            if (e != cudaSuccess) return e;
            // 32-bit hand indices inside the kernel: batches of 2^31 hands or more go in pieces
            constexpr int64_t kPiece = (int64_t)1 << 30;
            for (int64_t off = 0; off < n; off += kPiece, count_launch())
                eval_batch_fast_kernel<NC><<<sms, 1024, smem, st>>>(c2 + off, n - off < kPiece ? n - off : kPiece, out + off, T,
                                                                     1u);
        }
    } else {
        int blocks = (int)((n + 255) / 256);
        if (blocks > sms * 8) blocks = sms * 8;
        count_launch(); eval_batch_kernel<NC, false><<<blocks, 256, 1024, st>>>(c2, n, out, T);
    }
    return cudaGetLastError();
}

cudaError_t launch_eval_batch(const DevTables &T, const uint8_t *cards, int n_cards, int64_t n, uint16_t *out,
                              int sms, bool strict, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    switch (n_cards) {
        case 5: return launch_eval_nc<5>(T, cards, n, out, sms, strict, st);
        case 6: return launch_eval_nc<6>(T, cards, n, out, sms, strict, st);
        case 7: return launch_eval_nc<7>(T, cards, n, out, sms, strict, st);
    }
    return cudaErrorInvalidValue;
}

// =====================================================================================================================
// Whole-table showdown (GameBoard.determine_winner, GameBoard.py:204-229): one warp per hand, one lane per seat (<= 22,
// GameBoard.py:12).  Each lane ranks its seat's seven cards (two hole cards + the shared five-card board) with the strict
// routine, the warp reduces the best (lowest Treys rank: the correct winner) and the worst (highest: what the reference's
// max() at GameBoard.py:210 picks) over the active seats, and writes the ranks plus two winner bit masks per hand.
// =====================================================================================================================
__global__ void __launch_bounds__(256)
table_showdown_kernel(const uint8_t *__restrict__ hole, const uint8_t *__restrict__ board, const uint8_t *__restrict__ active,
                      int64_t hands, int seats, uint16_t *__restrict__ ranks, uint32_t *__restrict__ winners, DevTables T) {
    __shared__ uint32_t s_cardkey[64];
    if (threadIdx.x < 64) s_cardkey[threadIdx.x] = threadIdx.x < 52 ? c_rank_key[threadIdx.x % 13] : 0u;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const NfHash h7 = T.nf[2];
    const int64_t n_warps = (int64_t)gridDim.x * 8;
    for (int64_t hnd = (int64_t)blockIdx.x * 8 + warp; hnd < hands; hnd += n_warps) {
        uint32_t r = 0;
        bool act = false;
        if ((int)lane < seats) {
            const uint8_t *b = board + 5 * hnd, *hc = hole + 2 * (hnd * seats + lane);
            uint2 w;
            w.x = hc[0] | (hc[1] << 8) | (b[0] << 16) | ((uint32_t)b[1] << 24);
            w.y = b[2] | (b[3] << 8) | (b[4] << 16) | 0xFF000000u;
            r = eval_row<7>(w, s_cardkey, T.flush, h7);            // 0 = malformed seat (bad card or duplicate)
            act = (active == nullptr || active[hnd * seats + lane] != 0) && r != 0;
            ranks[hnd * seats + lane] = (uint16_t)r;
        }
        const uint32_t best = __reduce_min_sync(0xFFFFFFFFu, act ? r : 0xFFFFu);
        const uint32_t worst = __reduce_max_sync(0xFFFFFFFFu, act ? r : 0u);
        const uint32_t mb = __ballot_sync(0xFFFFFFFFu, act && r == best), mw = __ballot_sync(0xFFFFFFFFu, act && r == worst);
        if (lane == 0) { winners[2 * hnd] = mb; winners[2 * hnd + 1] = mw; }
    }
}

cudaError_t launch_table_showdown(const DevTables &T, const uint8_t *hole, const uint8_t *board, const uint8_t *active,
                                  int64_t hands, int seats, uint16_t *ranks, uint32_t *winners, int sms, cudaStream_t st) {
    if (hands == 0) return cudaSuccess;
    int64_t blocks = (hands + 7) / 8;
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    count_launch(); table_showdown_kernel<<<(int)blocks, 256, 0, st>>>(hole, board, active, hands, seats, ranks, winners, T);
    return cudaGetLastError();
}

// =====================================================================================================================
// Literal category ranker over rows of n_cards (2..9) cards, duplicates allowed.
// =====================================================================================================================
__global__ void __launch_bounds__(256)
category_batch_kernel(const uint8_t *__restrict__ cards, int n_cards, int row_stride, int64_t n,
                      uint8_t *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint8_t *row = cards + i * row_stride;
        LitHand h = {0, 0, 0, 0, 0};
        bool bad = false;
        for (int k = 0; k < n_cards; ++k) {
            uint32_t c = row[k];
            bad |= c > 51u;
            lit_add_card(h, c > 51u ? 0u : c);
        }
        out[i] = bad ? 0xFFu : (uint8_t)lit_category(h);
    }
}

cudaError_t launch_category_batch(const uint8_t *cards, int n_cards, int row_stride, int64_t n, uint8_t *out,
                                  int sms, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    int blocks = (int)((n + 255) / 256);
    if (blocks > sms * 8) blocks = sms * 8;
    count_launch(); category_batch_kernel<<<blocks, 256, 0, st>>>(cards, n_cards, row_stride, n, out);
    return cudaGetLastError();
}

// =====================================================================================================================
// Exhaustive sweep over all C(52,7) = 133,784,560 hands (BASELINE.json configs[1]), enumerated in-kernel in colex
// order: hand number i <-> c0 < c1 < ... < c6 with i = sum_k C(c_k, k+1).  Each thread unranks the start of its
// chunk once and then steps to the successor combination.  All three tables live in shared memory next to a
// per-CTA rank histogram.
// =====================================================================================================================
constexpr int64_t kHands7 = 133784560;
constexpr int kChunk = 512;  // hands per thread

__device__ __forceinline__ uint32_t binom_dev(const uint32_t *bt, int n, int k) { return bt[n * 8 + k]; }

__global__ void __launch_bounds__(1024, 1)
exhaustive7_kernel(unsigned long long *__restrict__ hist, unsigned long long *__restrict__ sums, DevTables T) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t *s_cardkey = reinterpret_cast<uint32_t *>(smem);           // 64
    uint32_t *s_binom = s_cardkey + 64;                                 // [53][8]
    uint32_t *s_hist = s_binom + 53 * 8;                                // [7464]
    uint16_t *s_flush = reinterpret_cast<uint16_t *>(s_hist + 7464);
    uint16_t *s_disp = s_flush + kFlushEntries;
    uint16_t *s_val = s_disp + T.disp_entries[2];
    if (threadIdx.x < 64) s_cardkey[threadIdx.x] = threadIdx.x < 52 ? c_rank_key[threadIdx.x % 13] : 0u;
    if (threadIdx.x < 53) {  // Pascal row per thread
        uint64_t v = 1;
        int nrow = threadIdx.x;
        for (int k = 0; k < 8; ++k) {
            s_binom[nrow * 8 + k] = k > nrow ? 0u : (uint32_t)v;
            v = v * (uint64_t)(nrow - k) / (uint64_t)(k + 1);
        }
    }
    for (uint32_t i = threadIdx.x; i < 7464; i += blockDim.x) s_hist[i] = 0;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(T.flush);
        uint4 *dst = reinterpret_cast<uint4 *>(s_flush);
        for (uint32_t i = threadIdx.x; i < kFlushEntries / 8; i += blockDim.x) dst[i] = src[i];
        src = reinterpret_cast<const uint4 *>(T.nf[2].disp); dst = reinterpret_cast<uint4 *>(s_disp);
        for (uint32_t i = threadIdx.x; i < T.disp_entries[2] / 8; i += blockDim.x) dst[i] = src[i];
        src = reinterpret_cast<const uint4 *>(T.nf[2].val); dst = reinterpret_cast<uint4 *>(s_val);
        for (uint32_t i = threadIdx.x; i < T.val_entries[2] / 8; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    NfHash h = T.nf[2];
    h.disp = s_disp;
    h.val = s_val;

    unsigned long long sum = 0, sumsq = 0;
    const int64_t n_chunks = (kHands7 + kChunk - 1) / kChunk;
    for (int64_t chunk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; chunk < n_chunks;
         chunk += (int64_t)gridDim.x * blockDim.x) {
        int64_t first = chunk * kChunk;
        int64_t last = first + kChunk < kHands7 ? first + kChunk : kHands7;
        // unrank `first` in the combinatorial number system
        int c[7];
        {
            uint32_t rem = (uint32_t)first;
            int hiC = 51;
            for (int k = 6; k >= 0; --k) {
                int v = hiC;
                while (binom_dev(s_binom, v, k + 1) > rem) --v;
                c[k] = v;
                rem -= binom_dev(s_binom, v, k + 1);
                hiC = v - 1;
            }
        }
        for (int64_t i = first; i < last; ++i) {
            uint32_t key = 0, lo = 0, hi = 0;
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                key += s_cardkey[c[k]];
                lo |= shl_clamp(1u, (uint32_t)c[k]);
                hi |= shl_clamp(1u, (uint32_t)c[k] - 32u);
            }
            uint32_t r = treys_rank(h, s_flush, key, ((uint64_t)hi << 32) | lo);
            atomicAdd(&s_hist[r], 1u);
            sum += r;
            sumsq += (unsigned long long)r * r;
            // successor in colex order: bump the lowest index that has room, reset everything below it
            if (c[0] + 1 < c[1]) { c[0]++; }
            else if (c[1] + 1 < c[2]) { c[1]++; c[0] = 0; }
            else if (c[2] + 1 < c[3]) { c[2]++; c[0] = 0; c[1] = 1; }
            else if (c[3] + 1 < c[4]) { c[3]++; c[0] = 0; c[1] = 1; c[2] = 2; }
            else if (c[4] + 1 < c[5]) { c[4]++; c[0] = 0; c[1] = 1; c[2] = 2; c[3] = 3; }
            else if (c[5] + 1 < c[6]) { c[5]++; c[0] = 0; c[1] = 1; c[2] = 2; c[3] = 3; c[4] = 4; }
            else { c[6]++; c[0] = 0; c[1] = 1; c[2] = 2; c[3] = 3; c[4] = 4; c[5] = 5; }
        }
    }
    // block reduction of the two checksums, then one global atomic each
    for (int off = 16; off > 0; off >>= 1) {
        sum += __shfl_down_sync(0xFFFFFFFFu, sum, off);
        sumsq += __shfl_down_sync(0xFFFFFFFFu, sumsq, off);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&sums[0], sum);
        atomicAdd(&sums[1], sumsq);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < 7463; i += blockDim.x)
        if (s_hist[i]) atomicAdd(&hist[i], (unsigned long long)s_hist[i]);
}

// Second form of the sweep.  A warp fixes the five highest cards c2 < ... < c6 (a 5-subset of 50 values, enumerated in
// colex order) and its lanes run over the C(c2, 2) pairs c0 < c1 below them, 32 per pass: the five-card part of the
// key / suit-counter word and of the card set is warp-uniform and carried from tuple to tuple (only c2 moves between most
// successive tuples), so a hand costs two conflict-free card gathers, one add and the two hash gathers.  A flush is only
// possible when the five high cards hold three of a suit (warp-uniform test); then the suit-counter pattern of the sum
// picks the lanes that go to the suited table.  Chunks of 16 tuples are dealt to the CTAs round-robin (3.6 % spread of
// work between SMs) and claimed by the warps of a CTA from a shared-memory counter, heaviest (= last in colex) first.
constexpr uint32_t kTuples5 = 2118760;            // C(50, 5)
constexpr uint32_t kTup = 16;
constexpr uint32_t kTupChunks = (kTuples5 + kTup - 1) / kTup;
constexpr uint32_t kSweepPairs = 1128;            // C(48, 2): c1 <= 46 < c2 <= 47
constexpr uint32_t kSweepHead = kTab2Bytes + 64 + 50 * 8 * 4 + 2272 + 16 + 7464 * 4;

__global__ void __launch_bounds__(1024, 1)
exhaustive7_tuple_kernel(unsigned long long *__restrict__ hist, unsigned long long *__restrict__ sums, DevTables T) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t *s_tab = smem;                                                    // [52 rows of 256 B][32 lane words]
    uint32_t *s_slow = reinterpret_cast<uint32_t *>(s_tab + kTab2Bytes);      // 512 bits: 7-card suit-counter pattern holds a flush
    uint32_t *s_binom = s_slow + 16;                                          // [50][8]
    uint16_t *s_pairs = reinterpret_cast<uint16_t *>(s_binom + 50 * 8);       // colex pairs, c0 | c1 << 8
    uint32_t *s_ctr = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(s_pairs) + 2272);
    uint32_t *s_hist = s_ctr + 4;                                             // [7464]
    uint16_t *s_flush = reinterpret_cast<uint16_t *>(s_hist + 7464);
    uint16_t *s_disp = s_flush + kFlushEntries;
    uint16_t *s_val = s_disp + T.disp_entries[2];
    for (uint32_t i = threadIdx.x; i < 52 * 32; i += blockDim.x) {
        const uint32_t c = i >> 5, su = c / 13;
        *reinterpret_cast<uint32_t *>(s_tab + c * 256 + (i & 31u) * 4) =
            c_rank_key[c % 13] | (su < 3 ? 1u << (23 + 3 * su) : 0u);
    }
    if (threadIdx.x < 16) {
        uint32_t word = 0;
        for (uint32_t b = 0; b < 32; ++b) {
            const uint32_t sv = threadIdx.x * 32 + b, c0 = sv & 7u, c1 = (sv >> 3) & 7u, c2 = sv >> 6;
            if (c0 >= 5 || c1 >= 5 || c2 >= 5 || c0 + c1 + c2 <= 2u) word |= 1u << b;
        }
        s_slow[threadIdx.x] = word;
    }
    if (threadIdx.x < 50) {  // Pascal row per thread
        uint64_t v = 1;
        const int nrow = threadIdx.x;
        for (int k = 0; k < 8; ++k) {
            s_binom[nrow * 8 + k] = k > nrow ? 0u : (uint32_t)v;
            v = v * (uint64_t)(nrow - k) / (uint64_t)(k + 1);
        }
    }
    for (uint32_t i = threadIdx.x; i < kSweepPairs; i += blockDim.x) s_pairs[i] = T.pairs[i];
    if (threadIdx.x == 0) s_ctr[0] = 0;
    for (uint32_t i = threadIdx.x; i < 7464; i += blockDim.x) s_hist[i] = 0;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(T.flush);
        uint4 *dst = reinterpret_cast<uint4 *>(s_flush);
        for (uint32_t i = threadIdx.x; i < kFlushEntries / 8; i += blockDim.x) dst[i] = src[i];
        src = reinterpret_cast<const uint4 *>(T.nf[2].disp); dst = reinterpret_cast<uint4 *>(s_disp);
        for (uint32_t i = threadIdx.x; i < T.disp_entries[2] / 8; i += blockDim.x) dst[i] = src[i];
        src = reinterpret_cast<const uint4 *>(T.nf[2].val); dst = reinterpret_cast<uint4 *>(s_val);
        for (uint32_t i = threadIdx.x; i < T.val_entries[2] / 8; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const NfHash &h = T.nf[2];
    const uint32_t lane = threadIdx.x & 31, lane4 = lane * 4;
    const uint8_t *tab_lane = s_tab + lane4;
    auto word_of = [&](uint32_t c) { return *reinterpret_cast<const uint32_t *>(tab_lane + c * 256); };
    // this CTA's chunks: blockIdx.x, blockIdx.x + gridDim.x, ...
    const uint32_t nk = blockIdx.x < kTupChunks ? (kTupChunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;

    unsigned long long sum = 0, sumsq = 0;
    for (;;) {
        uint32_t k = 0;
        if (lane == 0) k = atomicAdd(&s_ctr[0], 1u);
        k = __shfl_sync(0xFFFFFFFFu, k, 0);
        if (k >= nk) break;
        uint32_t q = (blockIdx.x + gridDim.x * (nk - 1 - k)) * kTup;
        const uint32_t qend = q + kTup < kTuples5 ? q + kTup : kTuples5;
        // unrank q among the 5-subsets of {0..49} (combinatorial number system)
        int d[5];
        {
            uint32_t rem = q;
            int hi = 49;
#pragma unroll
            for (int j = 4; j >= 0; --j) {
                int v = hi;
                while (s_binom[v * 8 + j + 1] > rem) --v;
                d[j] = v;
                rem -= s_binom[v * 8 + j + 1];
                hi = v - 1;
            }
        }
        bool fresh = true;
        uint32_t hi4_sum = 0;
        uint64_t hi4_set = 0;
        for (; q < qend; ++q) {
            if (fresh) {  // c3..c6 changed
                hi4_sum = word_of(d[1] + 2) + word_of(d[2] + 2) + word_of(d[3] + 2) + word_of(d[4] + 2);
                hi4_set = card_bit(d[1] + 2) | card_bit(d[2] + 2) | card_bit(d[3] + 2) | card_bit(d[4] + 2);
            }
            const uint32_t c2 = d[0] + 2;
            const uint32_t hi_sum = hi4_sum + word_of(c2);
            const uint64_t hi_set = hi4_set | card_bit(c2);
            // three of a suit among the five high cards?  (spades are the cards not counted)
            const uint32_t sc = hi_sum >> 23, a0 = sc & 7u, a1 = (sc >> 3) & 7u, a2 = sc >> 6;
            const bool flush_possible = a0 >= 3 || a1 >= 3 || a2 >= 3 || a0 + a1 + a2 <= 2u;
            const uint32_t npairs = (c2 * (c2 - 1)) >> 1;
            for (uint32_t p = lane; p < npairs; p += 32) {
                const uint32_t pr = s_pairs[p];
                // byte offsets lane * 4 | card << 8 of the two low cards
                const uint32_t w0 = *reinterpret_cast<const uint32_t *>(s_tab + __byte_perm(pr, lane4, 0x5504u));
                const uint32_t w1 = *reinterpret_cast<const uint32_t *>(s_tab + __byte_perm(pr, lane4, 0x5514u));
                const uint32_t tot = hi_sum + w0 + w1;
                const uint32_t key = tot & 0x7FFFFFu;
                const uint32_t b = (key * h.mul_b) >> h.bshift;
                const uint32_t slot = (((key * h.mul_s) >> h.sshift) + s_disp[b]) & 0xFFFFu;
                uint32_t r = s_val[slot];
                if (flush_possible) {
                    const uint32_t s = tot >> 23;
                    if ((s_slow[s >> 5] >> (s & 31u)) & 1u)
                        r = s_flush[flush_mask(hi_set | card_bit(pr & 0xFFu) | card_bit(pr >> 8), 5)];
                }
                atomicAdd(&s_hist[r], 1u);
                sum += r;
                sumsq += (unsigned long long)(r * r);
            }
            // successor in colex order
            if (d[0] + 1 < d[1]) { d[0]++; fresh = false; }
            else {
                fresh = true;
                if (d[1] + 1 < d[2]) { d[1]++; d[0] = 0; }
                else if (d[2] + 1 < d[3]) { d[2]++; d[0] = 0; d[1] = 1; }
                else if (d[3] + 1 < d[4]) { d[3]++; d[0] = 0; d[1] = 1; d[2] = 2; }
                else { d[4]++; d[0] = 0; d[1] = 1; d[2] = 2; d[3] = 3; }
            }
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        sum += __shfl_down_sync(0xFFFFFFFFu, sum, off);
        sumsq += __shfl_down_sync(0xFFFFFFFFu, sumsq, off);
    }
    if (lane == 0) {
        atomicAdd(&sums[0], sum);
        atomicAdd(&sums[1], sumsq);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < 7463; i += blockDim.x)
        if (s_hist[i]) atomicAdd(&hist[i], (unsigned long long)s_hist[i]);
}

cudaError_t launch_exhaustive7(const DevTables &T, uint64_t *hist, uint64_t *sums, int sms, cudaStream_t st, int variant) {
    cudaError_t e = cudaMemsetAsync(hist, 0, 7463 * sizeof(uint64_t), st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(sums, 0, 2 * sizeof(uint64_t), st);
    if (e != cudaSuccess) return e;
    const bool chunked = variant == 1;
    if (!chunked) {
        if (T.val_entries[2] != 65536u) return cudaErrorInvalidValue;   // the kernel masks slots with 0xFFFF
        const size_t smem2 = kSweepHead + 2 * (size_t)(kFlushEntries + T.disp_entries[2] + T.val_entries[2]);
        e = cudaFuncSetAttribute(exhaustive7_tuple_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        if (e != cudaSuccess) return e;
        count_launch(); exhaustive7_tuple_kernel<<<sms, 1024, smem2, st>>>(reinterpret_cast<unsigned long long *>(hist),
                                                          reinterpret_cast<unsigned long long *>(sums), T);
        return cudaGetLastError();
    }
    size_t smem = (64 + 53 * 8 + 7464) * 4 + 2 * (size_t)(kFlushEntries + T.disp_entries[2] + T.val_entries[2]);
    e = cudaFuncSetAttribute(exhaustive7_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    count_launch(); exhaustive7_kernel<<<sms, 1024, smem, st>>>(reinterpret_cast<unsigned long long *>(hist),
                                               reinterpret_cast<unsigned long long *>(sums), T);
    return cudaGetLastError();
}

// =====================================================================================================================
// Shared-memory gather micro-benchmark: the measured "lookup roofline" (SURVEY.md section 8d).  Every lane of a full
// grid issues `iters` x 4 independent chains of dependent uint16 gathers over a table of the production footprint.
// =====================================================================================================================
__global__ void __launch_bounds__(1024)
smem_gather_kernel(int table_entries, int iters, int pattern, uint32_t *__restrict__ sink) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint16_t *tab = reinterpret_cast<uint16_t *>(smem);
    uint32_t x = 0x9E3779B9u * (blockIdx.x + 1);
    for (int i = threadIdx.x; i < table_entries; i += blockDim.x) {
        uint32_t v = (uint32_t)i * 2654435761u + x;
        v ^= v >> 15; v *= 0x2C1B3C6Du; v ^= v >> 12;
        tab[i] = (uint16_t)v;
    }
    __syncthreads();
    const uint32_t mask = (uint32_t)table_entries - 1;  // power of two
    const uint32_t lane = threadIdx.x & 31;
    uint32_t a0 = threadIdx.x * 7919u + 1, a1 = a0 * 31u + 7, a2 = a1 * 31u + 7, a3 = a2 * 31u + 7;
    uint32_t acc = 0;
    if (pattern == 0) {  // uniformly random indices: bank conflicts as they fall
        for (int it = 0; it < iters; ++it) {
            uint32_t v0 = tab[a0 & mask], v1 = tab[a1 & mask], v2 = tab[a2 & mask], v3 = tab[a3 & mask];
            a0 = a0 * 1664525u + v0; a1 = a1 * 1664525u + v1; a2 = a2 * 1664525u + v2; a3 = a3 * 1664525u + v3;
            acc += v0 ^ v1 ^ v2 ^ v3;
        }
    } else {             // conflict-free: warp-uniform random base, lane-strided 32-bit words
        uint32_t w = (threadIdx.x >> 5) * 7919u + 1;
        uint32_t b0 = w, b1 = w * 31u + 7, b2 = b1 * 31u + 7, b3 = b2 * 31u + 7;
        for (int it = 0; it < iters; ++it) {
            uint32_t v0 = tab[((b0 << 6) + 2 * lane) & mask], v1 = tab[((b1 << 6) + 2 * lane) & mask];
            uint32_t v2 = tab[((b2 << 6) + 2 * lane) & mask], v3 = tab[((b3 << 6) + 2 * lane) & mask];
            // keep the next base warp-uniform but data dependent
            b0 = b0 * 1664525u + __shfl_sync(0xFFFFFFFFu, v0, 0);
            b1 = b1 * 1664525u + __shfl_sync(0xFFFFFFFFu, v1, 0);
            b2 = b2 * 1664525u + __shfl_sync(0xFFFFFFFFu, v2, 0);
            b3 = b3 * 1664525u + __shfl_sync(0xFFFFFFFFu, v3, 0);
            acc += v0 ^ v1 ^ v2 ^ v3;
        }
    }
    if (acc == 0x12345678u) sink[0] = acc;  // never true in practice; keeps the loop alive
}

cudaError_t launch_smem_gather(int table_bytes, int iters, int pattern, int blocks, int threads, cudaStream_t st) {
    static uint32_t *sink = nullptr;
    if (!sink) {
        cudaError_t e = cudaMalloc(&sink, 4);
        if (e != cudaSuccess) return e;
    }
    cudaError_t e = cudaFuncSetAttribute(smem_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, table_bytes);
    if (e != cudaSuccess) return e;
    count_launch(); smem_gather_kernel<<<blocks, threads, table_bytes, st>>>(table_bytes / 2, iters, pattern, sink);
    return cudaGetLastError();
}

// =====================================================================================================================
// Integer issue micro-benchmark: eight independent chains per thread, each step one compare (ALU pipe) and one predicated
// multiply-add (FMA pipe) -- the instruction pair the HP kernels are made of.  mix 0: compare + multiply-add pairs;
// mix 1: multiply-adds only (FMA pipe); mix 2: add / xor pairs only (ALU pipe).  Returns nothing; timed by the caller.
// =====================================================================================================================
__global__ void __launch_bounds__(1024, 1)
int_issue_kernel(int iters, int mix, uint32_t one, uint32_t *__restrict__ sink) {
    uint32_t a[8], t = threadIdx.x * 2654435761u + blockIdx.x;
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = t + k * 977u;
    if (mix == 0) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %0, %1;\n\t@p mad.lo.u32 %0, %1, %2, %0;\n\t}"
                             : "+r"(a[k]) : "r"(a[(k + 1) & 7]), "r"(one));
        }
    } else if (mix == 1) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(one), "r"(t));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(one), "r"(t));
            }
        }
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[k]) : "r"(t));          // IADD3 then LOP3: cannot be merged
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[k]) : "r"(a[(k + 1) & 7]));
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc ^= a[k];
    if (acc == 0x12345678u) sink[0] = acc;
}

cudaError_t launch_int_issue(int iters, int mix, int blocks, int threads, cudaStream_t st) {
    static uint32_t *sink = nullptr;
    if (!sink) {
        cudaError_t e = cudaMalloc(&sink, 4);
        if (e != cudaSuccess) return e;
    }
    count_launch(); int_issue_kernel<<<blocks, threads, 0, st>>>(iters, mix, 1u, sink);
    return cudaGetLastError();
}

}  // namespace holdem
