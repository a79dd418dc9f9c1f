// kernels_hshp.cu -- hand strength (HandCalculations.py:157-176) and the plain-enumeration hand potential kernel
// (HandCalculations.py:178-226) for both rank semantics and flop/turn/river boards.  sm_100a only.
//
// hs_kernel:        one warp per situation; lanes stride over the C(47,2)/C(46,2)/C(45,2) opponent holdings; the
//                   ahead/tied/behind counts are reduced with __reduce_add_sync.
// hp_direct_kernel: one CTA per situation; one warp per opponent holding, lanes stride over the runouts; one opponent
//                   evaluation per (holding, runout) pair exactly like the reference's nested loop (:193-218), but
//                   our own final rank is computed once per runout instead of once per pair (:210).  This is the
//                   independent second implementation the deduplicating flop / turn kernels (kernels_hp_flop.cu,
//                   kernels_hp_turn.cu) are checked against, and the production path for HOLDEM_MODE_REFERENCE (all
//                   streets) and for Treys-mode river rows (HS only).
#include <cstdlib>
#include "device_common.cuh"
#include "launch.h"

namespace holdem {

namespace {

constexpr int kModeTreys = 0;
constexpr int kModeReference = 1;

struct Situation {
    uint32_t hole[2];
    uint32_t board[5];
    int nb;
    bool ok;
    uint64_t known;  // hole | board card set
    uint64_t bset;   // board card set
};

__device__ __forceinline__ Situation parse_situation(const uint8_t *sit, int64_t idx) {
    uint2 w = __ldg(reinterpret_cast<const uint2 *>(sit) + idx);
    Situation s;
    uint32_t b[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        b[k] = (w.x >> (8 * k)) & 0xFFu;
        b[4 + k] = (w.y >> (8 * k)) & 0xFFu;
    }
    s.hole[0] = b[0];
    s.hole[1] = b[1];
    s.nb = 0;
    bool ok = b[0] < 52u && b[1] < 52u;
    bool ended = false;
    uint64_t bset = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        uint32_t c = b[2 + k];
        s.board[k] = c;
        if (c == 0xFFu) ended = true;
        else {
            ok = ok && !ended && c < 52u;
            s.nb += 1;
            bset |= card_bit(c < 52u ? c : 63u);
        }
    }
    uint64_t known = bset | card_bit(b[0] < 52u ? b[0] : 63u) | card_bit(b[1] < 52u ? b[1] : 63u);
    ok = ok && s.nb >= 3 && (__popcll(known) == s.nb + 2);
    s.ok = ok;
    s.known = known;
    s.bset = bset;
    return s;
}

// Ascending list of the cards NOT in `excluded`, built by one warp: lane l owns cards l and l + 32.
__device__ __forceinline__ int warp_list_complement(uint64_t excluded, uint8_t *list, uint32_t lane) {
    bool in0 = !((excluded >> lane) & 1ull);
    bool in1 = (lane + 32 < 52) && !((excluded >> (lane + 32)) & 1ull);
    uint32_t m0 = __ballot_sync(0xFFFFFFFFu, in0), m1 = __ballot_sync(0xFFFFFFFFu, in1);
    uint32_t lt = (1u << lane) - 1u;
    if (in0) list[__popc(m0 & lt)] = (uint8_t)lane;
    if (in1) list[__popc(m0) + __popc(m1 & lt)] = (uint8_t)(lane + 32);
    return __popc(m0) + __popc(m1);
}

__device__ __forceinline__ LitHand lit_from_cards(const uint32_t (&c)[5], int n) {
    LitHand h = {0, 0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 5; ++k)
        if (k < n) lit_add_card(h, c[k]);
    return h;
}

__device__ __forceinline__ LitHand lit_pair(uint32_t a, uint32_t b) {
    LitHand h = {0, 0, 0, 0, 0};
    lit_add_card(h, a);
    lit_add_card(h, b);
    return h;
}

}  // namespace

// =====================================================================================================================
// HS: one warp per situation
// =====================================================================================================================
constexpr int kHsWarps = 8;

__global__ void __launch_bounds__(kHsWarps * 32)
hs_kernel(const uint8_t *__restrict__ sit, int64_t n, int mode, uint32_t *__restrict__ out, DevTables T) {
    __shared__ uint32_t s_cardkey[64];
    __shared__ uint8_t s_unseen[kHsWarps][64];
    if (threadIdx.x < 64) s_cardkey[threadIdx.x] = threadIdx.x < 52 ? c_rank_key[threadIdx.x % 13] : 0u;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *unseen = s_unseen[warp];
    const int64_t n_warps = (int64_t)gridDim.x * kHsWarps;
    for (int64_t idx = (int64_t)blockIdx.x * kHsWarps + warp; idx < n; idx += n_warps) {
        Situation s = parse_situation(sit, idx);
        if (!s.ok) {
            if (lane < 3) out[idx * 3 + lane] = 0xFFFFFFFFu;
            continue;
        }
        __syncwarp();
        int m = warp_list_complement(s.known, unseen, lane);
        __syncwarp();
        int n_opp = m * (m - 1) / 2;
        uint32_t ahead = 0, tied = 0, total = 0;
        if (mode == kModeTreys) {
            const NfHash &h = T.nf[s.nb - 3];
            uint32_t kb = 0;
#pragma unroll
            for (int k = 0; k < 5; ++k) kb += k < s.nb ? s_cardkey[s.board[k]] : 0u;
            uint32_t ours = treys_rank(h, T.flush, kb + s_cardkey[s.hole[0]] + s_cardkey[s.hole[1]], s.known);
            for (int p = lane; p < n_opp; p += 32) {
                uint32_t pr = __ldg(T.pairs + p);
                uint32_t o1 = unseen[pr & 0xFFu], o2 = unseen[pr >> 8];
                uint32_t r = treys_rank(h, T.flush, kb + s_cardkey[o1] + s_cardkey[o2],
                                        s.bset | card_bit(o1) | card_bit(o2));
                ahead += ours < r;  // lower Treys rank is the better hand
                tied += ours == r;
                total += 1;
            }
        } else {
            LitHand lb = lit_from_cards(s.board, s.nb);
            uint32_t ours = lit_category(lit_merge(lb, lit_pair(s.hole[0], s.hole[1])));
            for (int p = lane; p < n_opp; p += 32) {
                uint32_t pr = __ldg(T.pairs + p);
                uint32_t o1 = unseen[pr & 0xFFu], o2 = unseen[pr >> 8];
                uint32_t r = lit_category(lit_merge(lb, lit_pair(o1, o2)));
                ahead += ours > r;  // HandCalculations.py:168
                tied += ours == r;
                total += 1;
            }
        }
        ahead = __reduce_add_sync(0xFFFFFFFFu, ahead);
        tied = __reduce_add_sync(0xFFFFFFFFu, tied);
        total = __reduce_add_sync(0xFFFFFFFFu, total);
        if (lane == 0) {
            out[idx * 3 + 0] = ahead;
            out[idx * 3 + 1] = tied;
            out[idx * 3 + 2] = total - ahead - tied;
        }
    }
}

cudaError_t launch_hs_batch(const DevTables &T, const uint8_t *sit, int64_t n, int mode, uint32_t *out, int sms,
                            cudaStream_t st, int variant) {
    if (n == 0) return cudaSuccess;
    if (mode != kModeTreys && variant == 0) return launch_hs_ref5(sit, n, out, sms, st);
    if (mode == kModeTreys && variant == 0) return launch_hs_treys_v4(T, sit, n, out, sms, st);
    int64_t blocks = (n + kHsWarps - 1) / kHsWarps;
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    count_launch(); hs_kernel<<<(int)blocks, kHsWarps * 32, 0, st>>>(sit, n, mode, out, T);
    return cudaGetLastError();
}

// =====================================================================================================================
// HP, plain enumeration: one CTA per situation
// =====================================================================================================================
constexpr int kHpThreads = 256;
constexpr int kHpWarps = kHpThreads / 32;
constexpr int kMaxOpp = 1081;   // C(47,2)
constexpr int kMaxRun = 1176;   // C(49,2), literal flop runouts

struct HpSmem {
    uint32_t cardkey[64];
    uint32_t cnt[12];           // HP[9], HPTotal[3]
    uint8_t unseen[64];         // deck \ (hole u board)
    uint8_t rdeck[64];          // literal mode: deck \ board
    // per opponent holding
    uint64_t p_cards[kMaxOpp];
    uint64_t p_rc[kMaxOpp];     // literal: rank nibbles
    uint32_t p_key[kMaxOpp];    // treys: additive key; literal: sc | rmask << 16
    uint8_t p_cls[kMaxOpp + 3]; // 0 ahead, 1 tied, 2 behind on the current board
    // per runout
    uint64_t q_cards[kMaxRun];
    uint64_t q_rc[kMaxRun];
    uint32_t q_key[kMaxRun];
    uint16_t q_our[kMaxRun];    // our final rank / category on board + runout
};

template <int MODE>
__global__ void __launch_bounds__(kHpThreads)
hp_direct_kernel(const uint8_t *__restrict__ sit, int64_t n, uint32_t *__restrict__ out, DevTables T,
                 int *__restrict__ run_flag, int skip_boards) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    HpSmem &S = *reinterpret_cast<HpSmem *>(smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // second pass after the flop kernel: nothing to do unless that kernel saw turn/river rows
    if (run_flag != nullptr && (*run_flag & 2) == 0) return;
    if (tid < 64) S.cardkey[tid] = tid < 52 ? c_rank_key[tid % 13] : 0u;

    for (int64_t idx = blockIdx.x; idx < n; idx += gridDim.x) {
        __syncthreads();  // previous situation fully consumed (and cardkey visible on the first pass)
        Situation s = parse_situation(sit, idx);
        if (skip_boards && s.board[4] == 0xFFu) continue;  // 3- and 4-card boards are done by the flop / turn kernels
        if (!s.ok) {
            if (tid == 0 && run_flag != nullptr) atomicOr(run_flag, 4);   // bit 2: a malformed row was marked
            if (tid < 16) out[idx * 16 + tid] = 0xFFFFFFFFu;
            continue;
        }
        if (tid < 12) S.cnt[tid] = 0;
        const int m = 52 - 2 - s.nb;
        if (warp == 0) warp_list_complement(s.known, S.unseen, lane);
        if (warp == 1) warp_list_complement(s.bset, S.rdeck, lane);
        __syncthreads();
        const int n_opp = m * (m - 1) / 2;
        const NfHash &h7 = T.nf[2];

        // -------- per-holding and per-runout tables ---------------------------------------------------------------
        int n_run;  // runouts enumerated per opponent holding (before the disjointness filter of treys mode)
        uint32_t kb = 0;
        LitHand lb = {0, 0, 0, 0, 0};
        uint32_t c0 = 0, c1 = 0, c2 = 0;
        if (MODE == kModeTreys) {
            const int run_cards = 5 - s.nb;
            n_run = run_cards == 2 ? n_opp : (run_cards == 1 ? m : 0);
            const NfHash &hn = T.nf[s.nb - 3];
#pragma unroll
            for (int k = 0; k < 5; ++k) kb += k < s.nb ? S.cardkey[s.board[k]] : 0u;
            const uint32_t khole = S.cardkey[s.hole[0]] + S.cardkey[s.hole[1]];
            const uint32_t ours_now = treys_rank(hn, T.flush, kb + khole, s.known);
            for (int p = tid; p < n_opp; p += kHpThreads) {
                uint32_t pr = __ldg(T.pairs + p);
                uint32_t a = S.unseen[pr & 0xFFu], b = S.unseen[pr >> 8];
                uint32_t key = S.cardkey[a] + S.cardkey[b];
                uint64_t cs = card_bit(a) | card_bit(b);
                S.p_key[p] = key;
                S.p_cards[p] = cs;
                uint32_t r = treys_rank(hn, T.flush, kb + key, s.bset | cs);
                uint32_t cls = ours_now < r ? 0u : (ours_now == r ? 1u : 2u);  // lower Treys rank = better
                S.p_cls[p] = (uint8_t)cls;
                c0 += cls == 0; c1 += cls == 1; c2 += cls == 2;
                if (run_cards == 2) {  // the same pair seen as a turn+river runout: our final 7-card rank
                    S.q_key[p] = key;
                    S.q_cards[p] = cs;
                    S.q_our[p] = (uint16_t)treys_rank(h7, T.flush, kb + khole + key, s.known | cs);
                }
            }
            if (run_cards == 1)
                for (int q = tid; q < m; q += kHpThreads) {
                    uint32_t a = S.unseen[q];
                    S.q_key[q] = S.cardkey[a];
                    S.q_cards[q] = card_bit(a);
                    S.q_our[q] = (uint16_t)treys_rank(h7, T.flush, kb + khole + S.cardkey[a], s.known | card_bit(a));
                }
        } else {
            const int mr = 52 - s.nb;
            n_run = mr * (mr - 1) / 2;
            lb = lit_from_cards(s.board, s.nb);
            const LitHand lours = lit_merge(lb, lit_pair(s.hole[0], s.hole[1]));
            const uint32_t ours_now = lit_category(lours);
            for (int p = tid; p < n_opp; p += kHpThreads) {
                uint32_t pr = __ldg(T.pairs + p);
                LitHand lp = lit_pair(S.unseen[pr & 0xFFu], S.unseen[pr >> 8]);
                S.p_cards[p] = lp.cards;
                S.p_rc[p] = lp.rc;
                S.p_key[p] = lp.sc | (lp.rmask << 16);
                uint32_t r = lit_category(lit_merge(lb, lp));
                uint32_t cls = ours_now > r ? 0u : (ours_now == r ? 1u : 2u);  // :196-201
                S.p_cls[p] = (uint8_t)cls;
                c0 += cls == 0; c1 += cls == 1; c2 += cls == 2;
            }
            for (int q = tid; q < n_run; q += kHpThreads) {
                uint32_t pr = __ldg(T.pairs + q);
                LitHand lq = lit_pair(S.rdeck[pr & 0xFFu], S.rdeck[pr >> 8]);  // :152-154: deck minus board only
                S.q_cards[q] = lq.cards;
                S.q_rc[q] = lq.rc;
                S.q_key[q] = lq.sc | (lq.rmask << 16);
                S.q_our[q] = (uint16_t)lit_category(lit_merge(lours, lq));      // :210, may repeat our hole cards
            }
        }
        c0 = __reduce_add_sync(0xFFFFFFFFu, c0);
        c1 = __reduce_add_sync(0xFFFFFFFFu, c1);
        c2 = __reduce_add_sync(0xFFFFFFFFu, c2);
        if (lane == 0) { atomicAdd(&S.cnt[9], c0); atomicAdd(&S.cnt[10], c1); atomicAdd(&S.cnt[11], c2); }  // :203
        __syncthreads();

        // -------- main enumeration: warp per opponent holding, lanes over runouts (:193-218) --------------------------
        uint32_t hp[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int p = warp; p < n_opp && n_run > 0; p += kHpWarps) {
            const uint32_t cls = S.p_cls[p];
            const uint64_t pcs = S.p_cards[p];
            uint32_t ahead = 0, tied = 0, behind = 0;
            if (MODE == kModeTreys) {
                const uint32_t kp = kb + S.p_key[p];
                const uint64_t bp = s.bset | pcs;
                for (int q = lane; q < n_run; q += 32) {
                    uint64_t qcs = S.q_cards[q];
                    if (qcs & pcs) continue;  // the runout must come from the cards neither player holds
                    uint32_t r = treys_rank(h7, T.flush, kp + S.q_key[q], bp | qcs);
                    uint32_t o = S.q_our[q];
                    ahead += o < r; tied += o == r; behind += o > r;
                }
            } else {
                LitHand lp;
                lp.cards = pcs; lp.rc = S.p_rc[p]; lp.dup = 0;
                lp.sc = S.p_key[p] & 0xFFFFu; lp.rmask = S.p_key[p] >> 16;
                const LitHand lbp = lit_merge(lb, lp);
                for (int q = lane; q < n_run; q += 32) {
                    LitHand lq;
                    lq.cards = S.q_cards[q]; lq.rc = S.q_rc[q]; lq.dup = 0;
                    lq.sc = S.q_key[q] & 0xFFFFu; lq.rmask = S.q_key[q] >> 16;
                    uint32_t r = lit_category(lit_merge(lbp, lq));  // :211, multiset when the runout repeats a card
                    uint32_t o = S.q_our[q];
                    ahead += o > r; tied += o == r; behind += o < r;
                }
            }
            if (cls == 0) { hp[0] += ahead; hp[1] += tied; hp[2] += behind; }
            else if (cls == 1) { hp[3] += ahead; hp[4] += tied; hp[5] += behind; }
            else { hp[6] += ahead; hp[7] += tied; hp[8] += behind; }
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            uint32_t v = __reduce_add_sync(0xFFFFFFFFu, hp[k]);
            if (lane == 0 && v) atomicAdd(&S.cnt[k], v);
        }
        __syncthreads();
        if (tid < 16) {
            uint32_t v;
            if (tid < 3) v = S.cnt[9 + tid];          // ahead, tied, behind == HPTotal (same loop as HandStrength)
            else if (tid < 12) v = S.cnt[tid - 3];    // HP[3][3]
            else if (tid < 15) v = S.cnt[tid - 3];    // HPTotal[3]
            else v = (uint32_t)s.nb;
            out[idx * 16 + tid] = v;
        }
    }
}

cudaError_t launch_hp_direct(const DevTables &T, const uint8_t *sit, int64_t n, int mode, uint32_t *out, int sms,
                             int *run_flag, int skip_boards, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const size_t smem = sizeof(HpSmem);
    cudaError_t e;
    int64_t blocks = n < (int64_t)sms * 4 ? n : (int64_t)sms * 4;
    if (mode == kModeTreys) {
        e = cudaFuncSetAttribute(hp_direct_kernel<kModeTreys>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        count_launch(); hp_direct_kernel<kModeTreys><<<(int)blocks, kHpThreads, smem, st>>>(sit, n, out, T, run_flag, skip_boards);
    } else {
        e = cudaFuncSetAttribute(hp_direct_kernel<kModeReference>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        count_launch(); hp_direct_kernel<kModeReference><<<(int)blocks, kHpThreads, smem, st>>>(sit, n, out, T, run_flag, skip_boards);
    }
    return cudaGetLastError();
}

}  // namespace holdem
