// kernels_hs4.cu -- hand strength (HandCalculations.py:157-176 with rank = Treys evaluate) for flop, turn and river boards
// from rank pairs: the hand-strength step of the rank-multiset HP kernels as a kernel of its own.  sm_100a only.
//
//   ahead / tied / behind over the C(m,2) opponent holdings, m = 47 / 46 / 45 unseen cards.
//   A holding that does not make a flush with the board is ranked by its RANK PAIR alone, so the counts are
//       sum over the 91 rank pairs pp of  n(pp) x cmp(ours, NF(board + pp)),        n(pp) = unseen pairs of those ranks,
//   with NF(board + pp) one row of a direct-indexed table (board_opp5 / board4_opp6 / own_our7, tables.h) and `ours` one
//   entry of own_rank5 / own6_rank6 / own_our7 -- no hash lookups.  Holdings that DO make a flush with the board (both cards
//   in a suit with >= 3 board cards, one card in a suit with >= 4, any holding on a board of five of a suit) are moved from
//   the class of their rank pair to the class of their suited-table rank: <= 45 / 9 x 36 + 36 / 45 + 10 x 35 + 1 items.
// One warp per situation, lanes over rank pairs (3 each); 8 B in, 12 B out per situation.
#include "device_common.cuh"
#include "launch.h"
#include "multiset_common.cuh"

namespace holdem {

namespace {

using namespace ms;

constexpr int kHs4Warps = 8;

// number of unseen cards of rank r: four minus the known ones (bit 13 s + r of the card set)
__device__ __forceinline__ uint32_t unseen_of_rank(uint64_t known, uint32_t r) {
    return 4u - (uint32_t)__popcll((known >> r) & 0x0000008004002001ull);
}

__device__ __forceinline__ uint32_t cls_of(uint32_t ours, uint32_t theirs) {   // 0 ahead, 1 tied, 2 behind (lower rank wins)
    return ours < theirs ? 0u : (ours == theirs ? 1u : 2u);
}

}  // namespace

// kHpRows: second pass of holdem_hp_batch for the rows the flop / turn kernels left (five-card boards: hand strength only).
// Writes 16-word HP rows (ahead tied behind | nine zeros | HPTotal | 5), skips every other row, exits at once unless bit 1 of
// *run_flag is set and raises bit 2 when it marks a malformed row.
template <bool kHpRows>
__global__ void __launch_bounds__(kHs4Warps * 32)
hs_treys_v4_kernel(const uint8_t *__restrict__ sit, int64_t n, uint32_t *__restrict__ out, DevTables T, int *__restrict__ run_flag) {
    if (kHpRows && run_flag != nullptr && (*run_flag & 2) == 0) return;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n_warps = (int64_t)gridDim.x * kHs4Warps;
    // the lane's three rank pairs
    uint32_t px[3], py[3];
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        const uint32_t pp = min(lane + 32u * p, 90u);
        uint32_t y = 0;
        while (tri(y + 1) <= pp) ++y;
        px[p] = pp - tri(y);
        py[p] = y;
    }
    for (int64_t idx = (int64_t)blockIdx.x * kHs4Warps + warp; idx < n; idx += n_warps) {
        const uint2 w = __ldg(reinterpret_cast<const uint2 *>(sit) + idx);
        uint32_t cb[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) { cb[k] = (w.x >> (8 * k)) & 0xFFu; cb[4 + k] = (w.y >> (8 * k)) & 0xFFu; }
        if (kHpRows && cb[6] == 0xFFu) continue;          // a flop or turn row: done by the HP kernels of those streets
        // board length: 3, 4 or 5 contiguous cards
        const uint32_t nb = kHpRows ? 5u : (cb[5] == 0xFFu ? 3u : (cb[6] == 0xFFu ? 4u : 5u));
        bool ok = true;
        uint64_t known = 0, bset = 0;
        uint32_t bsuit = 0;
        uint32_t rk[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            const uint32_t c = cb[k];
            const bool used = (uint32_t)k < nb + 2;
            ok = ok && (used ? c < 52u : c == 0xFFu);
            const uint32_t cc = (used && c < 52u) ? c : 0u;
            rk[k] = used ? card_rank(cc) : 0u;
            if (used) {
                known |= card_bit(cc);
                if (k >= 2) { bset |= card_bit(cc); bsuit += 1u << (4 * card_suit(cc)); }
            }
        }
        ok = ok && (uint32_t)__popcll(known) == nb + 2;
        if (!ok) {
            if (kHpRows) {
                if (lane < 16) out[idx * 16 + lane] = 0xFFFFFFFFu;
                if (lane == 0 && run_flag != nullptr) atomicOr(run_flag, 4);
            } else if (lane < 3) out[idx * 3 + lane] = 0xFFFFFFFFu;
            continue;
        }
        // ---- table rows: lane 0 sorts the ranks and forms the multiset indices ------------------------------------------------
        uint32_t row_idx = 0, ours = 0;
        if (lane == 0) {
            // board ranks ascending (insertion network over up to five values)
            uint32_t b[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) b[k] = (uint32_t)k < nb ? rk[2 + k] : 99u;
            cswap(b[0], b[1]); cswap(b[3], b[4]); cswap(b[2], b[4]); cswap(b[2], b[3]); cswap(b[0], b[3]);   // 9-comparator network
            cswap(b[0], b[2]); cswap(b[1], b[4]); cswap(b[1], b[3]); cswap(b[1], b[2]);
            uint32_t bi = b[0] + tri(b[1]) + (b[2] + 2) * (b[2] + 1) * b[2] / 6;
            if (nb >= 4) bi += (b[3] + 3) * (b[3] + 2) * (b[3] + 1) * b[3] / 24;
            if (nb == 5) bi += (b[4] + 4) * (b[4] + 3) * (b[4] + 2) * (b[4] + 1) * b[4] / 120;
            row_idx = bi;
            const uint32_t fm = flush_mask(known, 5);
            if (fm) ours = __ldg(T.flush + fm);
            else if (nb == 5) {
                const uint32_t h0 = min(rk[0], rk[1]), h1 = max(rk[0], rk[1]);
                ours = __ldg(T.own_our7 + (size_t)bi * 96 + tri(h1) + h0);
            } else {
                // hole + board multiset: insert the two hole ranks into the sorted board
                uint32_t a[7];
                a[0] = min(rk[0], rk[1]); a[1] = max(rk[0], rk[1]);
#pragma unroll
                for (int k = 0; k < 4; ++k) a[2 + k] = b[k];      // b[3] = 99 on the flop: it stays last and is not used
                cswap(a[1], a[2]); cswap(a[2], a[3]); cswap(a[3], a[4]); cswap(a[4], a[5]);
                cswap(a[0], a[1]); cswap(a[1], a[2]); cswap(a[2], a[3]); cswap(a[3], a[4]);
                uint32_t mi = a[0] + tri(a[1]) + (a[2] + 2) * (a[2] + 1) * a[2] / 6 + (a[3] + 3) * (a[3] + 2) * (a[3] + 1) * a[3] / 24 +
                              (a[4] + 4) * (a[4] + 3) * (a[4] + 2) * (a[4] + 1) * a[4] / 120;
                if (nb == 3) ours = __ldg(T.own_rank5 + mi);
                else {
                    mi += (a[5] + 5) * (a[5] + 4) * (a[5] + 3) * (a[5] + 2) * (a[5] + 1) * a[5] / 720;
                    ours = __ldg(T.own6_rank6 + mi);
                }
            }
        }
        row_idx = __shfl_sync(0xFFFFFFFFu, row_idx, 0);
        ours = __shfl_sync(0xFFFFFFFFu, ours, 0);
        const uint16_t *row = nb == 3 ? T.board_opp5 + (size_t)row_idx * 96
                                      : (nb == 4 ? T.board4_opp6 + (size_t)row_idx * 96 : T.own_our7 + (size_t)row_idx * 96);
        // ---- rank-pair sum -------------------------------------------------------------------------------------------------------
        uint32_t acc[3] = {0, 0, 0};        // ahead, tied, behind (wraps mod 2^32 while holdings are moved between classes)
        uint32_t cl[3];
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            const uint32_t pp = lane + 32u * p;
            const uint32_t v = pp < 91u ? (uint32_t)__ldg(row + pp) : 0u;
            const uint32_t ux = unseen_of_rank(known, px[p]), uy = unseen_of_rank(known, py[p]);
            const uint32_t cnt = (pp < 91u && v != 0u) ? (px[p] != py[p] ? ux * uy : (ux * (ux - 1u)) >> 1) : 0u;
            cl[p] = v ? cls_of(ours, v) : 3u;
#pragma unroll
            for (uint32_t c = 0; c < 3; ++c) acc[c] += cl[p] == c ? cnt : 0u;
        }
        // ---- holdings that make a flush with the board ---------------------------------------------------------------------------
        const uint32_t m3 = (bsuit + 0x5555u) & 0x8888u;            // a suit with >= 3 board cards (at most one)
        if (m3) {
            const uint32_t s = (uint32_t)(__ffs(m3) - 1) >> 2, bs = (bsuit >> (4 * s)) & 15u;
            const uint32_t bmask = (uint32_t)(bset >> (13 * s)) & 0x1FFFu;
            const uint32_t smask = ~(uint32_t)(known >> (13 * s)) & 0x1FFFu;     // unseen ranks of the suit
            const uint32_t nns = (50u - nb) - (uint32_t)__popc(smask);            // unseen cards outside the suit
            // both cards in s: the 78 rank pairs x < y, three per lane (pairs of equal rank do not exist inside a suit)
#pragma unroll
            for (int p = 0; p < 3; ++p) {
                const uint32_t pp = lane + 32u * p;
                const bool both = pp < 91u && px[p] != py[p] && ((smask >> px[p]) & 1u) && ((smask >> py[p]) & 1u);
                if (both) {
                    const uint32_t ct = cls_of(ours, __ldg(T.flush + (bmask | (1u << px[p]) | (1u << py[p]))));
#pragma unroll
                    for (uint32_t c = 0; c < 3; ++c) acc[c] += (uint32_t)(ct == c) - (uint32_t)(cl[p] == c);
                }
            }
            if (bs >= 4) {
                // one card a in s (lane = its rank), the other card c outside: all of them play the flush of a
                if (lane < 13 && ((smask >> lane) & 1u)) {
                    const uint32_t ct = cls_of(ours, __ldg(T.flush + (bmask | (1u << lane))));
#pragma unroll
                    for (uint32_t c = 0; c < 3; ++c) acc[c] += ct == c ? nns : 0u;
                    for (uint32_t y = 0; y < 13; ++y) {           // minus the classes their rank pairs (a, y) were counted in
                        const uint32_t wy = unseen_of_rank(known, y) - ((smask >> y) & 1u);
                        const uint32_t v = __ldg(row + tri(max(lane, y)) + min(lane, y));
                        const uint32_t cr = v ? cls_of(ours, v) : 3u;
#pragma unroll
                        for (uint32_t c = 0; c < 3; ++c) acc[c] -= cr == c ? wy : 0u;
                    }
                }
            }
            if (bs == 5) {
                // the board itself is a flush: holdings without a card of s play it
                const uint32_t ct = cls_of(ours, __ldg(T.flush + bmask));
                if (lane == 0) {
#pragma unroll
                    for (uint32_t c = 0; c < 3; ++c) acc[c] += ct == c ? (nns * (nns - 1u)) >> 1 : 0u;
                }
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    const uint32_t pp = lane + 32u * p;
                    const uint32_t wx = unseen_of_rank(known, px[p]) - ((smask >> px[p]) & 1u);
                    const uint32_t wy = unseen_of_rank(known, py[p]) - ((smask >> py[p]) & 1u);
                    const uint32_t wn = pp < 91u ? (px[p] != py[p] ? wx * wy : (wx * (wx - 1u)) >> 1) : 0u;
#pragma unroll
                    for (uint32_t c = 0; c < 3; ++c) acc[c] -= cl[p] == c ? wn : 0u;
                }
            }
        }
#pragma unroll
        for (uint32_t c = 0; c < 3; ++c) acc[c] = __reduce_add_sync(0xFFFFFFFFu, acc[c]);
        if (kHpRows) {
            if (lane < 16) {
                const uint32_t k = lane < 3 ? lane : (lane >= 12 && lane < 15 ? lane - 12 : 3u);
                out[idx * 16 + lane] = k < 3 ? (k == 0 ? acc[0] : (k == 1 ? acc[1] : acc[2])) : (lane == 15 ? 5u : 0u);
            }
        } else if (lane < 3) out[idx * 3 + lane] = lane == 0 ? acc[0] : (lane == 1 ? acc[1] : acc[2]);
    }
}

cudaError_t launch_hs_treys_v4(const DevTables &T, const uint8_t *sit, int64_t n, uint32_t *out, int sms, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    int64_t blocks = (n + kHs4Warps - 1) / kHs4Warps;
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    count_launch(); hs_treys_v4_kernel<false><<<(int)blocks, kHs4Warps * 32, 0, st>>>(sit, n, out, T, nullptr);
    return cudaGetLastError();
}

cudaError_t launch_hp_river_rows(const DevTables &T, const uint8_t *sit, int64_t n, uint32_t *out, int sms, int *run_flag,
                                 cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    int64_t blocks = (n + kHs4Warps - 1) / kHs4Warps;
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    count_launch(); hs_treys_v4_kernel<true><<<(int)blocks, kHs4Warps * 32, 0, st>>>(sit, n, out, T, run_flag);
    return cudaGetLastError();
}

}  // namespace holdem
