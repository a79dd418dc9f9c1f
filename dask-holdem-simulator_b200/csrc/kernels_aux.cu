// kernels_aux.cu -- small elementwise kernels either side of the HS/HP kernels and the
// content hash the multi-GPU parity checks use:
//   * rows_hash_kernel        position-dependent 64-bit hash of a [n, words] uint32 buffer (bench.py / tests: equal across
//                             GPU counts and kernel variants means equal rows, unlike a sum of counts)
//   * decide_kernel           Decision.decision_based_on_blind_position (Decision.py:4-48) over whole batches, either from
//                             (hand strength, ppot, npot) doubles or straight from HS/HP count rows (HandCalculations.py:175
//                             for HS, the normalised potentials of holdem_b200/api.py for PPot/NPot)
//   * pack_rank_major_kernel  utils.card_to_int wire encoding (utils.py:82-85: rank * 4 + suit) -> situation rows in the path
//                             encoding (HandCalculations.py:53-54: suit * 13 + rank), validated
// All are HBM-bound streaming kernels: one thread per row, coalesced 16-byte loads where the row allows it.
#include "launch.h"

namespace holdem {

namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t x) {     // splitmix64 finaliser
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}

// hash = sum over rows r of mix64(fnv1a64(row words, seeded with first_row + r)): a sum, so it does not depend on which
// thread or which GPU hashed a row, and it is position-dependent through the seed.
__global__ void __launch_bounds__(256)
rows_hash_kernel(const uint32_t *__restrict__ rows, int64_t n, int words, int64_t first_row,
                 unsigned long long *__restrict__ out) {
    unsigned long long acc = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
        uint64_t h = 0xCBF29CE484222325ull ^ mix64((uint64_t)(first_row + r) + 0x9E3779B97F4A7C15ull);
        const uint32_t *p = rows + r * words;
        if ((words & 3) == 0 && (reinterpret_cast<uintptr_t>(rows) & 15) == 0) {
            for (int w = 0; w < words; w += 4) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p + w));
                h = (h ^ v.x) * 0x100000001B3ull; h = (h ^ v.y) * 0x100000001B3ull;
                h = (h ^ v.z) * 0x100000001B3ull; h = (h ^ v.w) * 0x100000001B3ull;
            }
        } else {
            for (int w = 0; w < words; ++w) h = (h ^ __ldg(p + w)) * 0x100000001B3ull;
        }
        acc += mix64(h);
    }
    for (int d = 16; d; d >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

// Decision.py:4-48.  Codes: 0 fold, 1 check, 2 call, 3 raise, -1 = the reference falls through and returns None
// (hand_strength > 0.8 and npot >= 0.3).  pos: 0 'SB', 1 'BB', 2 'NONE'.
__device__ __forceinline__ int decide(double hs, double ppot, double npot, int pos) {
    if (hs > 0.8) {                                             // :12-19
        if (npot < 0.3) return pos == 0 ? 1 : (pos == 1 ? 2 : 3);
        return -1;
    }
    if (hs > 0.5) {                                             // :22-36
        if (ppot > 0.6) return pos == 0 ? 1 : 2;
        return pos == 1 ? 1 : 0;
    }
    if (pos == 1) return ppot > 0.8 ? 2 : 1;                     // :39-48
    return 0;
}

__global__ void __launch_bounds__(256)
decide_kernel(const double *__restrict__ hs, const double *__restrict__ ppot, const double *__restrict__ npot,
              const uint8_t *__restrict__ pos, int pos_all, int64_t n, int8_t *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int p = pos ? (int)pos[i] : pos_all;
        out[i] = p > 2 ? (int8_t)-2 : (int8_t)decide(hs[i], ppot[i], npot[i], p);
    }
}

// From count rows: HS as HandCalculations.py:175 in double; PPot/NPot = the formula of :221-224 with every HP row divided by
// the number of runouts per opponent (api.normalised_potentials, double, same operation order); NaN -> 0 (nobody ahead /
// behind) exactly like policy.decide_from_counts.  Optionally writes the three doubles (hs_ppot_npot [n,3]).
__global__ void __launch_bounds__(256)
decide_from_counts_kernel(const uint32_t *__restrict__ rows, const uint8_t *__restrict__ pos, int pos_all, int64_t n,
                          int8_t *__restrict__ out, double *__restrict__ hs_ppot_npot) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint4 *r4 = reinterpret_cast<const uint4 *>(rows + i * 16);
        const uint4 a = __ldg(r4), b = __ldg(r4 + 1), c = __ldg(r4 + 2), d = __ldg(r4 + 3);
        const int p = pos ? (int)pos[i] : pos_all;
        if (d.w == 0xFFFFFFFFu || p > 2) {                      // malformed situation row / unknown position
            out[i] = (int8_t)-2;
            if (hs_ppot_npot) { hs_ppot_npot[3 * i] = hs_ppot_npot[3 * i + 1] = hs_ppot_npot[3 * i + 2] = __longlong_as_double(0x7FF8000000000000ll); }
            continue;
        }
        const double ahead = (double)a.x, tied = (double)a.y, behind = (double)a.z;
        const double hs = (ahead + tied / 2) / (ahead + tied + behind);
        // HP[3][3] row-major in words 3..11, HPTotal in 12..14
        const double h00 = (double)a.w, h01 = (double)b.x, h02 = (double)b.y, h10 = (double)b.z, h11 = (double)b.w,
                     h12 = (double)c.x, h20 = (double)c.y, h21 = (double)c.z, h22 = (double)c.w;
        const double t0 = (double)d.x, t1 = (double)d.y, t2 = (double)d.z;
        double runouts = ((h00 + h01 + h02) + (h10 + h11 + h12) + (h20 + h21 + h22)) / (t0 + t1 + t2);   // integers: any order is exact
        if (!(runouts > 0)) runouts = 1.0;
        double pp = (h20 + h21 / 2 + h10 / 2) / ((t2 + t1) * runouts);
        double np = (h02 + h12 / 2 + h01 / 2) / ((t0 + t1) * runouts);
        if (pp != pp) pp = 0.0;
        if (np != np) np = 0.0;
        out[i] = (int8_t)decide(hs, pp, np, p);
        if (hs_ppot_npot) { hs_ppot_npot[3 * i] = hs; hs_ppot_npot[3 * i + 1] = pp; hs_ppot_npot[3 * i + 2] = np; }
    }
}

// rank-major wire ints (int64, utils.py:82-85) -> uint8 [n,8] situation rows.  A value outside [0,52) becomes 0xFE, which
// every kernel rejects as a malformed row (0xFF stays "absent board card").
__global__ void __launch_bounds__(256)
pack_rank_major_kernel(const int64_t *__restrict__ hole, const int64_t *__restrict__ board, int n_board, int64_t n,
                       uint8_t *__restrict__ sit) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t c[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) c[k] = 0xFFu;
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            if (k < 2 + n_board) {
                const int64_t v = k < 2 ? hole[2 * i + k] : board[(int64_t)n_board * i + (k - 2)];
                c[k] = (v >= 0 && v < 52) ? (uint32_t)((v & 3) * 13 + (v >> 2)) : 0xFEu;
            }
        }
        uint2 w;
        w.x = c[0] | (c[1] << 8) | (c[2] << 16) | (c[3] << 24);
        w.y = c[4] | (c[5] << 8) | (c[6] << 16) | (c[7] << 24);
        reinterpret_cast<uint2 *>(sit)[i] = w;
    }
}

// Table layout (GameBoard.py:12 MAX_PLAYERS = 22; one board per hand, shared by its seats) -> situation rows, seat-major
// inside a hand so that the seats of one board sit next to each other in the batch.
__global__ void __launch_bounds__(256)
table_rows_kernel(const uint8_t *__restrict__ hole, const uint8_t *__restrict__ board, int64_t hands, int seats, int n_board,
                  uint8_t *__restrict__ sit) {
    const int64_t total = hands * seats, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t h = i / seats;
        uint32_t c[8];
        c[0] = hole[2 * i]; c[1] = hole[2 * i + 1];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const uint32_t b = board[5 * h + k];
            c[2 + k] = k < n_board ? (b == 0xFFu ? 0xFEu : b) : 0xFFu;      // an absent card inside the used part is malformed
        }
        c[7] = 0xFFu;
        uint2 w;
        w.x = c[0] | (c[1] << 8) | (c[2] << 16) | (c[3] << 24);
        w.y = c[4] | (c[5] << 8) | (c[6] << 16) | (c[7] << 24);
        reinterpret_cast<uint2 *>(sit)[i] = w;
    }
}

int grid_for(int64_t n, int sms) {
    int64_t blocks = (n + 255) / 256;
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace

cudaError_t launch_rows_hash(const uint32_t *rows, int64_t n, int words, int64_t first_row, uint64_t *out, int sms,
                             cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(uint64_t), st);
    if (e != cudaSuccess || n == 0) return e;
    count_launch(); rows_hash_kernel<<<grid_for(n, sms), 256, 0, st>>>(rows, n, words, first_row,
                                                                        reinterpret_cast<unsigned long long *>(out));
    return cudaGetLastError();
}

cudaError_t launch_decide(const double *hs, const double *ppot, const double *npot, const uint8_t *pos, int pos_all,
                          int64_t n, int8_t *out, int sms, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    count_launch(); decide_kernel<<<grid_for(n, sms), 256, 0, st>>>(hs, ppot, npot, pos, pos_all, n, out);
    return cudaGetLastError();
}

cudaError_t launch_decide_from_counts(const uint32_t *rows, const uint8_t *pos, int pos_all, int64_t n, int8_t *out,
                                      double *hs_ppot_npot, int sms, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    count_launch(); decide_from_counts_kernel<<<grid_for(n, sms), 256, 0, st>>>(rows, pos, pos_all, n, out, hs_ppot_npot);
    return cudaGetLastError();
}

cudaError_t launch_pack_rank_major(const int64_t *hole, const int64_t *board, int n_board, int64_t n, uint8_t *sit,
                                   int sms, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    count_launch(); pack_rank_major_kernel<<<grid_for(n, sms), 256, 0, st>>>(hole, board, n_board, n, sit);
    return cudaGetLastError();
}

cudaError_t launch_table_rows(const uint8_t *hole, const uint8_t *board, int64_t hands, int seats, int n_board,
                              uint8_t *sit, int sms, cudaStream_t st) {
    if (hands * seats == 0) return cudaSuccess;
    count_launch(); table_rows_kernel<<<grid_for(hands * seats, sms), 256, 0, st>>>(hole, board, hands, seats, n_board, sit);
    return cudaGetLastError();
}

}  // namespace holdem
