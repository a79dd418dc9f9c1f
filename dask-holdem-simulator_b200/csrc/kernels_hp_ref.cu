// kernels_hp_ref.cu -- hand strength + hand potential in HOLDEM_MODE_REFERENCE: the reference's loops exactly as
// written (HandCalculations.py:178-226 with the 9-category ranker of :45-82), flop / turn / river boards, one CTA per
// situation.  sm_100a only.  Bit-identical counts to hp_direct_kernel<reference> (cross-checked in the parity tests)
// at a fraction of the work.
//
// The reference visits every (opponent holding P, runout Q) with P a pair of the m = 47/46/45 cards unseen by us and
// Q ANY pair of deck \ board (:152-154) -- 1176/1128/1081 runouts, including pairs that repeat our hole cards or the
// opponent's own cards, so the ranked hands are card MULTISETS.  Split the runouts of a holding P = {u,v}:
//   regular   Q inside the unseen cards and disjoint from P: C(m-2,2) per P.  The opponent's hand board + P + Q depends
//             only on the 4-SET P u Q, our hand only on Q, so these are handled like the Treys-mode flop kernel
//             (kernels_hp_flop.cu): C(m,4) category evaluations, each serving its 6 splits as compare-and-count steps.
//             Categories come from a per-situation table NF[1820] over rank multisets of the four added cards (two
//             nibbles per entry: the category without and with a flush present, because the reference tests "straight"
//             before "flush") plus a suit/straight-flush test on the rare flush path.
//   irregular Q touches P or our hole cards: 4m-2 = 186/182/178 per P, Q = {e, x} with e one of u, v, h0, h1.  Lane = P, the
//             warp walks x together; the category comes from a second per-situation table NF2[pair of ranks][pair of
//             ranks] (91 x 91 nibble pairs) and the suit logic is done on multiplicities (a repeated card counts twice
//             towards the flush, and forbids the straight flush, exactly as the reference's Counter / set code does).
#include "device_common.cuh"
#include "launch.h"

namespace holdem {

namespace {

constexpr int kRThreads = 256;
constexpr int kRWarps = kRThreads / 32;
constexpr int kRPitch = 49;        // odd pitch
constexpr int kRRows = 92;         // real rows + zero rows for padding lanes
constexpr int kRXPitch = 96;
constexpr int kRNf = 1820;

struct RefSmem {
    uint32_t W[kRRows * kRPitch];      // [i][j], i<j<m: our category with runout (i,j) << 19 | one-hot class of (i,j)
    uint32_t X[4 * kRXPitch];          // [suit][i]: C(r_i+2,3) | rank bit of i if it has that suit << 16
    uint32_t I[48];                    // per position: r | C(r+1,2) << 4 | C(r+3,4) << 11 | suit << 22
    uint32_t cnt[12];                  // behind[3], ahead[3] (by class), class totals[3]
    uint32_t bmask[4];                 // board rank mask per suit
    uint8_t NF[kRNf + 4];              // category of board + 4-rank multiset: low nibble plain, high nibble with a flush
    uint8_t NF2[91 * 91 + 3];          // the same table indexed by two rank pairs [pid(r,r')][pid(r'',r''')], pid(a<=b) = b(b+1)/2 + a
    uint8_t ourH[2][48];               // our category with runout {hole card h, position x}
    uint8_t code[64];                  // unseen cards, rank-major code rank*4+suit, ascending
};

__device__ __forceinline__ uint32_t r_path_of_code(uint32_t code) { return (code & 3u) * 13u + (code >> 2); }
__device__ __forceinline__ uint32_t r_choose2(uint32_t n) { return n * (n - 1) / 2; }
__device__ __forceinline__ uint32_t r_choose3(uint32_t n) { return n * (n - 1) * (n - 2) / 6; }
__device__ __forceinline__ uint32_t r_choose4(uint32_t n) { return n * (n - 1) * (n - 2) * (n - 3) / 24; }

__device__ __forceinline__ void r_add_if_lt(uint32_t &acc, uint32_t wq, uint32_t thr, uint32_t wp, uint32_t one) {
    asm("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %1, %2;\n\t@p mad.lo.u32 %0, %3, %4, %0;\n\t}"
        : "+r"(acc) : "r"(wq), "r"(thr), "r"(wp), "r"(one));
}
__device__ __forceinline__ void r_add_if_ge(uint32_t &acc, uint32_t wq, uint32_t thr, uint32_t wp, uint32_t one) {
    asm("{\n\t.reg .pred p;\n\tsetp.ge.u32 p, %1, %2;\n\t@p mad.lo.u32 %0, %3, %4, %0;\n\t}"
        : "+r"(acc) : "r"(wq), "r"(thr), "r"(wp), "r"(one));
}

__device__ __forceinline__ LitHand r_one_card(uint32_t c) {
    LitHand h = {0, 0, 0, 0, 0};
    lit_add_card(h, c);
    return h;
}

}  // namespace

__global__ void __launch_bounds__(kRThreads)
hp_ref_kernel(const uint8_t *__restrict__ sit, int64_t n, uint32_t *__restrict__ out, DevTables T, uint32_t one,
              int *__restrict__ bad_flag) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    RefSmem &S = *reinterpret_cast<RefSmem *>(smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t i = tid; i < (uint32_t)(kRRows * kRPitch); i += kRThreads) S.W[i] = 0;   // pad rows stay zero
    for (uint32_t i = tid; i < (uint32_t)(4 * kRXPitch); i += kRThreads) S.X[i] = 0;

    for (int64_t idx = blockIdx.x; idx < n; idx += gridDim.x) {
        __syncthreads();
        // ---- parse ---------------------------------------------------------------------------------------------------
        uint2 w = __ldg(reinterpret_cast<const uint2 *>(sit) + idx);
        uint32_t cb[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) { cb[k] = (w.x >> (8 * k)) & 0xFFu; cb[4 + k] = (w.y >> (8 * k)) & 0xFFu; }
        int nb = 0;
        bool ok = cb[0] < 52u && cb[1] < 52u, ended = false;
        uint64_t known = card_bit(cb[0] < 52u ? cb[0] : 63u) | card_bit(cb[1] < 52u ? cb[1] : 63u), bset = 0;
        LitHand lb = {0, 0, 0, 0, 0};
        uint32_t scb = 0x4444u;   // board suit nibbles biased by 4: bit 3 set <=> count >= 4
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
                scb += 1u << (4 * card_suit(cc));
            }
        }
        known |= bset;
        ok = ok && nb >= 3 && __popcll(known) == nb + 2;
        if (!ok) {
            if (tid == 0 && bad_flag != nullptr) atomicOr(bad_flag, 4);   // bit 2: a malformed row was marked
            if (tid < 16) out[idx * 16 + tid] = 0xFFFFFFFFu;
            continue;
        }
        const uint32_t m = 50u - (uint32_t)nb;   // unseen cards
        const uint32_t n_pairs = m * (m - 1) / 2;
        const LitHand lours = lit_merge(lb, lit_merge(r_one_card(cb[0]), r_one_card(cb[1])));
        const uint32_t ours_now = lit_category(lours);
        if (tid < 12) S.cnt[tid] = 0;
        if (tid < 4) S.bmask[tid] = (uint32_t)(bset >> (13 * tid)) & 0x1FFFu;
        if (warp == 0) {   // unseen cards in (rank, suit) order
            const uint32_t p0 = r_path_of_code(lane), p1 = r_path_of_code(lane + 32);
            const bool in0 = !((known >> p0) & 1ull);
            const bool in1 = (lane + 32 < 52) && !((known >> p1) & 1ull);
            const uint32_t m0 = __ballot_sync(0xFFFFFFFFu, in0), m1 = __ballot_sync(0xFFFFFFFFu, in1);
            const uint32_t lt = (1u << lane) - 1u;
            if (in0) S.code[__popc(m0 & lt)] = (uint8_t)lane;
            if (in1) S.code[__popc(m0) + __popc(m1 & lt)] = (uint8_t)(lane + 32);
        }
        // NF[1820]: literal category of board + every multiset of four ranks, without suits (low nibble) and the value the
        // hand takes when a non-straight-flush flush is present (high nibble): quads, full house and straight outrank
        // the flush test in the reference's decision list (:68-74), everything else becomes 5.
        for (uint32_t e = tid; e < (uint32_t)kRNf; e += kRThreads) {
            const uint32_t rr = __ldg(T.nf4_ranks + e);
            LitHand h = lb;
            h.sc = 0; h.cards = 0; h.dup = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t r = (rr >> (4 * q)) & 15u;
                h.rc += 1ull << (4 * r);
                h.rmask |= 1u << r;
            }
            const uint32_t nfc = lit_category(h);
            const uint32_t with_flush = (nfc == 4u || nfc >= 6u) ? nfc : 5u;
            S.NF[e] = (uint8_t)(nfc | (with_flush << 4));
        }
        for (uint32_t e = tid; e < 91u * 91u; e += kRThreads) {   // pair-of-pairs view for the irregular runouts
            const uint32_t p2 = e / 91u, q2 = e - 91u * p2;
            uint32_t r4[4];
            {   // unrank pid = b(b+1)/2 + a
                uint32_t b = 0;
                while ((b + 1) * (b + 2) / 2 <= p2) ++b;
                r4[0] = p2 - b * (b + 1) / 2; r4[1] = b;
                b = 0;
                while ((b + 1) * (b + 2) / 2 <= q2) ++b;
                r4[2] = q2 - b * (b + 1) / 2; r4[3] = b;
            }
            LitHand h = lb;
            h.sc = 0; h.cards = 0; h.dup = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) { h.rc += 1ull << (4 * r4[q]); h.rmask |= 1u << r4[q]; }
            const uint32_t nfc = lit_category(h);
            S.NF2[e] = (uint8_t)(nfc | (((nfc == 4u || nfc >= 6u) ? nfc : 5u) << 4));
        }
        __syncthreads();
        {   // per-position and per-pair tables
            if (tid < m) {
                const uint32_t cd = S.code[tid], r = cd >> 2, su = cd & 3u;
                S.I[tid] = r | (r_choose2(r + 1) << 4) | (r_choose4(r + 3) << 11) | (su << 22);
#pragma unroll
                for (uint32_t s2 = 0; s2 < 4; ++s2)
                    S.X[s2 * kRXPitch + tid] = r_choose3(r + 2) | ((su == s2 ? (1u << r) : 0u) << 16);
                const LitHand lx = r_one_card(r_path_of_code(cd));
                S.ourH[0][tid] = (uint8_t)lit_category(lit_merge(lours, lit_merge(r_one_card(cb[0]), lx)));   // :210, h0 doubled
                S.ourH[1][tid] = (uint8_t)lit_category(lit_merge(lours, lit_merge(r_one_card(cb[1]), lx)));
            }
            uint32_t c0 = 0, c1 = 0, c2 = 0;
            for (uint32_t p = tid; p < n_pairs; p += kRThreads) {
                const uint32_t pr = __ldg(T.pairs + p);
                const uint32_t i = pr & 0xFFu, j = pr >> 8;
                LitHand lp = r_one_card(r_path_of_code(S.code[i]));
                lit_add_card(lp, r_path_of_code(S.code[j]));
                const uint32_t our7 = lit_category(lit_merge(lours, lp));
                const uint32_t opp = lit_category(lit_merge(lb, lp));
                const uint32_t cls = ours_now > opp ? 0u : (ours_now == opp ? 1u : 2u);   // :196-201
                c0 += cls == 0; c1 += cls == 1; c2 += cls == 2;
                S.W[i * kRPitch + j] = (our7 << 19) | (1u << (6 * cls));
            }
            c0 = __reduce_add_sync(0xFFFFFFFFu, c0);
            c1 = __reduce_add_sync(0xFFFFFFFFu, c1);
            c2 = __reduce_add_sync(0xFFFFFFFFu, c2);
            if (lane == 0) { atomicAdd(&S.cnt[6], c0); atomicAdd(&S.cnt[7], c1); atomicAdd(&S.cnt[8], c2); }   // :203
        }
        __syncthreads();

        // ---- regular runouts: 4-set walk (lane = (a,b,d), c walks b+1..d-1) ---------------------------------------------------
        const uint32_t *task_hdr = m == 47 ? T.flop_task_hdr : T.ref_task_hdr[47 - m - 1];
        const uint32_t *task_lanes = m == 47 ? T.flop_lanes : T.ref_lanes[47 - m - 1];
        const uint32_t n_tasks = m == 47 ? T.n_flop_tasks : T.n_ref_tasks[47 - m - 1];
        uint32_t acc[6] = {0, 0, 0, 0, 0, 0};      // packed 3 x 6-bit class counters: "our < opp" x3 partitions, "our > opp" x3
        uint32_t wide[6] = {0, 0, 0, 0, 0, 0};     // behind[class], ahead[class]
        uint32_t since = 0;
        for (uint32_t t = warp; t < n_tasks; t += kRWarps) {
            const uint32_t hdr = __ldg(task_hdr + t);
            const uint32_t trip = __ldg(task_lanes + t * 32 + lane);
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
            uint32_t wab = 0, wad = 0, wbd = 0, nf = 0, mask6 = 0;
            bool two_suits = false;   // river boards only: two suits hold 4 of the 8 fixed cards, c decides which one flushes
            const uint32_t *pac = S.W + 47 * kRPitch, *pbc = pac, *pcd = pac, *px = S.X + 48;   // zero words
            if (trip != 0xFFFFFFFFu) {
                const uint32_t a = trip & 0xFFu, b = (trip >> 8) & 0xFFu, d = (trip >> 16) & 0xFFu;
                const uint32_t ia = S.I[a], ib = S.I[b], id = S.I[d];
                nf = (ia & 15u) + ((ib >> 4) & 127u) + ((id >> 11) & 2047u);
                const uint32_t sc6 = scb + (1u << (4 * (ia >> 22))) + (1u << (4 * (ib >> 22))) + (1u << (4 * (id >> 22)));
                const uint32_t fl = sc6 & 0x8888u;   // suits holding >= 4 of the fixed cards (board + a, b, d)
                two_suits = (fl & (fl - 1u)) != 0;
                uint32_t sf = 0;
                if (fl) {
                    sf = (uint32_t)(__ffs(fl) - 1) >> 2;
                    const uint32_t *xr = S.X + sf * kRXPitch;
                    mask6 = S.bmask[sf] | ((xr[a] | xr[b] | xr[d]) >> 16);
                }
                wab = S.W[a * kRPitch + b];
                wad = S.W[a * kRPitch + d];
                wbd = S.W[b * kRPitch + d];
                const uint32_t c0i = b + 1 + k0;
                pac = S.W + a * kRPitch + c0i;
                pbc = S.W + b * kRPitch + c0i;
                pcd = S.W + c0i * kRPitch + d;
                px = S.X + sf * kRXPitch + c0i;
            }
#pragma unroll 2
            for (uint32_t k = 0; k < iters; ++k) {
                const uint32_t wac = pac[k], wbc = pbc[k], wcd = pcd[k * kRPitch], x = px[k];
                const uint32_t e = S.NF[nf + (x & 0xFFFFu)];
                uint32_t R = e & 15u;
                const uint32_t fm = mask6 | (x >> 16);
                if (__popc(fm) >= 5) R = cyclic_run5(fm) ? 8u : (e >> 4);   // all cards distinct here: :62-65, else :68-74
                if (two_suits) {   // rare: rank the nine cards arithmetically
                    LitHand h9 = lb;
                    lit_add_card(h9, r_path_of_code(S.code[trip & 0xFFu]));
                    lit_add_card(h9, r_path_of_code(S.code[(trip >> 8) & 0xFFu]));
                    lit_add_card(h9, r_path_of_code(S.code[(trip >> 16) & 0xFFu]));
                    lit_add_card(h9, r_path_of_code(S.code[((trip >> 8) & 0xFFu) + 1 + k0 + k]));
                    R = lit_category(h9);
                }
                const uint32_t lo = R << 19, hi = lo + (1u << 19);
                r_add_if_lt(acc[0], wcd, lo, wab, one);  r_add_if_ge(acc[3], wcd, hi, wab, one);   // P=ab Q=cd
                r_add_if_lt(acc[0], wab, lo, wcd, one);  r_add_if_ge(acc[3], wab, hi, wcd, one);   // P=cd Q=ab
                r_add_if_lt(acc[1], wbd, lo, wac, one);  r_add_if_ge(acc[4], wbd, hi, wac, one);   // P=ac Q=bd
                r_add_if_lt(acc[1], wac, lo, wbd, one);  r_add_if_ge(acc[4], wac, hi, wbd, one);   // P=bd Q=ac
                r_add_if_lt(acc[2], wbc, lo, wad, one);  r_add_if_ge(acc[5], wbc, hi, wad, one);   // P=ad Q=bc
                r_add_if_lt(acc[2], wad, lo, wbc, one);  r_add_if_ge(acc[5], wad, hi, wbc, one);   // P=bc Q=ad
            }
        }
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            wide[0] += acc[q] & 63u; wide[1] += (acc[q] >> 6) & 63u; wide[2] += (acc[q] >> 12) & 63u;
            wide[3] += acc[3 + q] & 63u; wide[4] += (acc[3 + q] >> 6) & 63u; wide[5] += (acc[3 + q] >> 12) & 63u;
        }

        // ---- irregular runouts: Q = {e, x} with e in {u, v, h0, h1} and x an unseen position, plus Q = {h0, h1}.  Lane = holding
        //      P = (u,v); the warp walks x in step, so everything about x is warp-uniform and everything about P is hoisted.
        //      The opponent's category comes from NF2[pid(ru,rv)][pid(re,rx)]; suit counts are multiplicities (a repeated
        //      card counts twice, :56), a straight flush needs the flush suit's cards pairwise distinct (:26-27).
        {
            const uint32_t h0 = cb[0], h1 = cb[1];
            const uint32_t rh0 = card_rank(h0), sh0 = card_suit(h0), rh1 = card_rank(h1), sh1 = card_suit(h1);
            const uint64_t bh0 = card_bit(h0), bh1 = card_bit(h1);
            const uint32_t our_hh = lit_category(lit_merge(lours, lit_merge(r_one_card(h0), r_one_card(h1))));
            const uint32_t scb3 = scb - 0x1111u;   // bias 3: bit 3 of a nibble set <=> that suit holds >= 5 cards
            const uint32_t half = (m + 1) / 2;
            const uint32_t n_ptasks = (n_pairs + 31) / 32;
            for (uint32_t t = warp; t < 2 * n_ptasks; t += kRWarps) {
                const uint32_t p = (t >> 1) * 32 + lane;
                const bool live = p < n_pairs;
                const uint32_t pr = __ldg(T.pairs + (live ? p : 0u));
                const uint32_t u = pr & 0xFFu, v = pr >> 8;
                const uint32_t cdu = S.code[u], cdv = S.code[v];
                const uint32_t ru = cdu >> 2, su = cdu & 3u, rv = cdv >> 2, sv = cdv & 3u;
                const uint64_t bu = card_bit(r_path_of_code(cdu)), bv = card_bit(r_path_of_code(cdv));
                const uint64_t cs_bp = bset | bu | bv;
                const uint32_t sc_bp = scb3 + (1u << (4 * su)) + (1u << (4 * sv));
                const uint32_t p2row = (rv * (rv + 1) / 2 + ru) * 91u;
                const uint32_t wuv = S.W[u * kRPitch + v];
                const uint32_t cls = (wuv & 63u) ? 0u : ((wuv & (63u << 6)) ? 1u : 2u);
                uint32_t behind = 0, ahead = 0;

                // one runout {e, x}: e has rank re / suit se / bit be and is (e_in_p) or is not one of the opponent's own cards
                auto item = [&](uint32_t re, uint32_t se, uint64_t be, bool e_in_p, uint32_t ours, bool valid,
                                uint32_t rx, uint32_t sx, uint64_t bx, bool x_in_p) {
                    const uint32_t lo_r = re < rx ? re : rx, hi_r = re < rx ? rx : re;
                    const uint32_t e8 = S.NF2[p2row + hi_r * (hi_r + 1) / 2 + lo_r];
                    const uint32_t fl = (sc_bp + (1u << (4 * se)) + (1u << (4 * sx))) & 0x8888u;
                    uint32_t opp = e8 & 15u;
                    if (fl) {
                        const uint32_t s = (uint32_t)(__ffs(fl) - 1) >> 2;
                        opp = e8 >> 4;
                        const bool repeated = (e_in_p && se == s) || (x_in_p && sx == s);
                        const uint32_t fm = (uint32_t)((cs_bp | be | bx) >> (13 * s)) & 0x1FFFu;
                        if (!repeated && cyclic_run5(fm)) opp = 8u;
                    }
                    if (valid) { behind += ours < opp; ahead += ours > opp; }
                };

                const uint32_t x_lo = (t & 1u) ? half : 0u, x_hi = (t & 1u) ? m : half;
                for (uint32_t x = x_lo; x < x_hi; ++x) {
                    const uint32_t cdx = S.code[x];
                    const uint32_t rx = cdx >> 2, sx = cdx & 3u;
                    const uint64_t bx = card_bit(r_path_of_code(cdx));
                    const bool x_in_p = x == u || x == v;
                    // {u, x}, x != u (x == v is the runout Q == P); our category is that of the plain pair runout
                    item(ru, su, bu, true, S.W[(x < u ? x : u) * kRPitch + (x < u ? u : x)] >> 19, live && x != u,
                         rx, sx, bx, x_in_p);
                    // {v, x}, x outside P
                    item(rv, sv, bv, true, S.W[(x < v ? x : v) * kRPitch + (x < v ? v : x)] >> 19, live && !x_in_p,
                         rx, sx, bx, x_in_p);
                    // {h0, x}, {h1, x}: the runout repeats one of OUR hole cards (never one of the opponent's)
                    item(rh0, sh0, bh0, false, S.ourH[0][x], live, rx, sx, bx, x_in_p);
                    item(rh1, sh1, bh1, false, S.ourH[1][x], live, rx, sx, bx, x_in_p);
                }
                if ((t & 1u) == 0)   // {h0, h1}: once per holding
                    item(rh0, sh0, bh0, false, our_hh, live, rh1, sh1, bh1, false);
                if (live) {
                    if (cls == 0) { wide[0] += behind; wide[3] += ahead; }
                    else if (cls == 1) { wide[1] += behind; wide[4] += ahead; }
                    else { wide[2] += behind; wide[5] += ahead; }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const uint32_t v = __reduce_add_sync(0xFFFFFFFFu, wide[q]);
            if (lane == 0 && v) atomicAdd(&S.cnt[q], v);
        }
        __syncthreads();
        if (tid < 16) {
            uint32_t v;
            const uint32_t runouts = (m + 2) * (m + 1) / 2;   // every pair of deck \ board (:152-154)
            if (tid < 3) v = S.cnt[6 + tid];
            else if (tid < 12) {
                const uint32_t cls = (tid - 3) / 3, o = (tid - 3) % 3;
                const uint32_t behind = S.cnt[cls], ahead = S.cnt[3 + cls], all = runouts * S.cnt[6 + cls];
                v = o == 0 ? ahead : (o == 2 ? behind : all - ahead - behind);   // HP[class][ahead|tied|behind] (:213-218)
            } else if (tid < 15) v = S.cnt[6 + tid - 12];
            else v = (uint32_t)nb;
            out[idx * 16 + tid] = v;
        }
    }
}

cudaError_t launch_hp_ref(const DevTables &T, const uint8_t *sit, int64_t n, uint32_t *out, int sms, int *bad_flag,
                          cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const size_t smem = sizeof(RefSmem);
    cudaError_t e = cudaFuncSetAttribute(hp_ref_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    static int bps = 0;
    if (bps == 0) {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, hp_ref_kernel, kRThreads, smem);
        if (e != cudaSuccess) return e;
        if (bps < 1) bps = 1;
    }
    int64_t blocks = (int64_t)sms * bps;
    if (blocks > n) blocks = n;
    count_launch(); hp_ref_kernel<<<(int)blocks, kRThreads, smem, st>>>(sit, n, out, T, 1u, bad_flag);
    return cudaGetLastError();
}

}  // namespace holdem
