// kernels_order.cu -- work order of the Treys-mode flop and turn kernels.
//
// Both kernels are instruction-cache sensitive (thousands of SASS instructions, several situation groups per SM in different code
// regions) and their cost per situation varies 3x with the board texture.  Two small kernels counting-sort the rows of a batch by
// (street, board texture, cards of our most frequent suit), heaviest class first; the situation groups then claim work indices of
// that order and look the rows up when the claim is made.  Results are stored by row number, so nothing else changes.  Measured
// on B200: flop kernel +8.7 % (mixed 65,536-situation workload), turn kernel +12 %.
//
// scratch (uint32): [0..31] flop class counts  ([29] turn-shaped rows, [30] river-shaped rows, [31] flop-shaped rows),
//                   [32..63] flop cursors, [64..95] turn class counts, [96..127] turn cursors,
//                   [128 .. 128+n) all rows sorted by flop class (rows of other streets last),
//                   [128+n .. 128+n+T) the turn-shaped rows sorted by turn class.
#include "device_common.cuh"
#include "launch.h"

namespace holdem {

namespace {

__device__ __forceinline__ uint32_t max4(uint32_t packed) {      // four 8-bit counters
    return max(max(packed & 0xFFu, (packed >> 8) & 0xFFu), max((packed >> 16) & 0xFFu, packed >> 24));
}

// flop: 4 x texture (0 monotone, 1 two-tone, 2 rainbow) + (5 - cards of our most frequent suit among hole + board); 12 = not a
// well-formed flop row (worked last by the flop kernel: rows of other streets are skipped there, malformed ones marked)
__device__ __forceinline__ uint32_t flop_class(uint2 w) {
    const uint32_t c[5] = {w.x & 0xFFu, (w.x >> 8) & 0xFFu, (w.x >> 16) & 0xFFu, w.x >> 24, w.y & 0xFFu};
    if (((w.y >> 8) & 0xFFu) != 0xFFu) return 12u;
    uint32_t hs = 0, bsu = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        if (c[k] >= 52u) return 12u;
        const uint32_t su = card_suit(c[k]);
        hs += 1u << (8 * su);
        if (k >= 2) bsu += 1u << (8 * su);
    }
    return 4u * (3u - max4(bsu)) + (5u - max4(hs));
}

// turn: 4 x texture (0 four of a suit, 1 three, 2 two + two, 3 two, 4 rainbow) + min(3, cards of the board's suit we lack);
// 20 = malformed turn-shaped row (marked by the turn kernel)
__device__ __forceinline__ uint32_t turn_class(uint2 w) {
    const uint32_t c[6] = {w.x & 0xFFu, (w.x >> 8) & 0xFFu, (w.x >> 16) & 0xFFu, w.x >> 24, w.y & 0xFFu, (w.y >> 8) & 0xFFu};
    uint32_t hs = 0, bsu = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        if (c[k] >= 52u) return 20u;
        const uint32_t su = card_suit(c[k]);
        hs += 1u << (8 * su);
        if (k >= 2) bsu += 1u << (8 * su);
    }
    const uint32_t bmax = max4(bsu), hmax = max4(hs);
    const bool two_two = bmax == 2 && ((bsu + 0x7E7E7E7Eu) & 0x80808080u) != 0 &&
                         __popc((bsu + 0x7E7E7E7Eu) & 0x80808080u) == 2;        // two suits with two board cards each
    const uint32_t t = bmax == 4 ? 0u : (bmax == 3 ? 1u : (bmax == 2 ? (two_two ? 2u : 3u) : 4u));
    return 4u * t + min(3u, bmax + 2u - hmax);
}

__device__ __forceinline__ bool turn_shaped(uint2 w) { return ((w.y >> 8) & 0xFFu) != 0xFFu && ((w.y >> 16) & 0xFFu) == 0xFFu; }

__global__ void __launch_bounds__(256)
street_order_count_kernel(const uint8_t *__restrict__ sit, int64_t n, uint32_t *__restrict__ scratch) {
    __shared__ uint32_t hist[64];
    if (threadIdx.x < 64) hist[threadIdx.x] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint2 w = __ldg(reinterpret_cast<const uint2 *>(sit) + i);
        atomicAdd(&hist[flop_class(w)], 1u);
        // rows by the kernel that owns them: flop-shaped (fourth and fifth board card absent), turn-shaped, river-shaped
        const bool no4 = ((w.y >> 8) & 0xFFu) == 0xFFu, no5 = ((w.y >> 16) & 0xFFu) == 0xFFu;
        atomicAdd(&hist[no5 ? (no4 ? 31 : 29) : 30], 1u);
        if (turn_shaped(w)) atomicAdd(&hist[32 + turn_class(w)], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 64 && hist[threadIdx.x])
        atomicAdd(scratch + (threadIdx.x < 32 ? threadIdx.x : 32 + threadIdx.x), hist[threadIdx.x]);     // [0..31] and [64..95]
}

// one class at a time through a shared-memory prefix table would need a second pass; the lanes of a warp that share a class
// share one atomic instead (match.any)
__global__ void __launch_bounds__(256)
street_order_scatter_kernel(const uint8_t *__restrict__ sit, int64_t n, uint32_t *__restrict__ scratch) {
    __shared__ uint32_t base[64];                      // [0..31] flop classes, [32..63] turn classes
    const uint32_t n_flop = scratch[31], n_turn = scratch[29];
    if (n_flop == 0 && n_turn == 0) return;            // neither kernel will look at the order
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int k = 0; k < 13; ++k) { base[k] = acc; acc += scratch[k]; }
        acc = 0;
        for (int k = 0; k < 21; ++k) { base[32 + k] = acc; acc += scratch[64 + k]; }
    }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    uint32_t *perm_flop = scratch + kOrderHeader, *perm_turn = scratch + kOrderHeader + n;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t b0 = (int64_t)blockIdx.x * blockDim.x; b0 < n; b0 += stride) {            // warp-uniform trip count
        const int64_t i = b0 + threadIdx.x;
        uint2 w = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
        if (i < n) w = __ldg(reinterpret_cast<const uint2 *>(sit) + i);
        if (n_flop != 0) {
            const uint32_t cls = i < n ? flop_class(w) : 31u;
            const uint32_t m = __match_any_sync(0xFFFFFFFFu, cls);
            const uint32_t leader = (uint32_t)(__ffs(m) - 1);
            uint32_t pos = 0;
            if (lane == leader && cls < 13u) pos = atomicAdd(scratch + 32 + cls, (uint32_t)__popc(m));
            pos = __shfl_sync(0xFFFFFFFFu, pos, leader);
            if (cls < 13u) perm_flop[base[cls] + pos + __popc(m & ((1u << lane) - 1u))] = (uint32_t)i;
        }
        if (n_turn != 0) {
            const uint32_t cls = (i < n && turn_shaped(w)) ? turn_class(w) : 31u;
            const uint32_t m = __match_any_sync(0xFFFFFFFFu, cls);
            const uint32_t leader = (uint32_t)(__ffs(m) - 1);
            uint32_t pos = 0;
            if (lane == leader && cls < 21u) pos = atomicAdd(scratch + 96 + cls, (uint32_t)__popc(m));
            pos = __shfl_sync(0xFFFFFFFFu, pos, leader);
            if (cls < 21u) perm_turn[base[32 + cls] + pos + __popc(m & ((1u << lane) - 1u))] = (uint32_t)i;
        }
    }
}

}  // namespace

size_t street_order_scratch_bytes(int64_t n) { return ((size_t)kOrderHeader + 2 * (size_t)n) * sizeof(uint32_t); }

cudaError_t launch_street_order(const uint8_t *sit, int64_t n, uint32_t *scratch, int sms, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(scratch, 0, kOrderHeader * sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    int64_t blocks = (n + 255) / 256;
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    count_launch(); street_order_count_kernel<<<(int)blocks, 256, 0, st>>>(sit, n, scratch);
    count_launch(); street_order_scatter_kernel<<<(int)blocks, 256, 0, st>>>(sit, n, scratch);
    return cudaGetLastError();
}

}  // namespace holdem
