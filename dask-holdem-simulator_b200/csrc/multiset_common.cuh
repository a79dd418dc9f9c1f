// multiset_common.cuh -- device helpers shared by the rank-multiset HS+HP kernels (kernels_hp_flop4.cu, kernels_hp_turn4.cu):
// named group barriers, mbarrier + cp.async.bulk wrappers, the folded suited-table index, and the compare / predicated
// multiply-add steps.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace holdem {
namespace ms {

__device__ __forceinline__ uint32_t path_of_code4(uint32_t code) { return (code & 3u) * 13u + (code >> 2); }
__device__ __forceinline__ uint32_t tri(uint32_t n) { return n * (n + 1) / 2; }          // C(n+1,2)
__device__ __forceinline__ uint32_t choose2u(uint32_t n) { return n * (n - 1) / 2; }    // n >= 1 or n == 0
__device__ __forceinline__ uint32_t fold(uint32_t m) { return m ^ (m >> 6); }           // bijection on 13-bit masks, XOR-linear
template <int kGroupThreads>
__device__ __forceinline__ void group_sync_n(uint32_t gid) {   // named barrier gid + 1 over one situation group
    asm volatile("bar.sync %0, %1;" ::"r"(gid + 1), "n"(kGroupThreads) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra WAIT_DONE;\n\tbra WAIT_LOOP;\n\tWAIT_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared through the async proxy (TMA engine); completion is counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void cswap(uint32_t &a, uint32_t &b) { const uint32_t lo = min(a, b), hi = max(a, b); a = lo; b = hi; }
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}

// One step of the flush part for the lane's runout: valid = step and runout share no card of the suit;
//   if (our < rf) aL += wa;  if (our > rf) aG += wa;  if (our < rn) sL += ws;  if (our > rn) sG += ws.
// The adds are multiply-adds by a kernel argument equal to 1 so that they issue on the FMA pipe beside the compares (ALU).
__device__ __forceinline__ void flush_step(uint32_t &aL, uint32_t &aG, uint32_t &sL, uint32_t &sG, uint32_t our,
                                           uint32_t rf, uint32_t rn, uint32_t dx, uint32_t qb2, uint32_t wa, uint32_t ws,
                                           uint32_t one) {
    asm("{\n\t.reg .pred v, p;\n\t.reg .b32 t;\n\t"
        "and.b32 t, %7, %8;\n\t"
        "setp.eq.u32 v, t, 0;\n\t"
        "setp.lt.and.u32 p, %4, %5, v;\n\t@p mad.lo.u32 %0, %9, %11, %0;\n\t"
        "setp.gt.and.u32 p, %4, %5, v;\n\t@p mad.lo.u32 %1, %9, %11, %1;\n\t"
        "setp.lt.and.u32 p, %4, %6, v;\n\t@p mad.lo.u32 %2, %10, %11, %2;\n\t"
        "setp.gt.and.u32 p, %4, %6, v;\n\t@p mad.lo.u32 %3, %10, %11, %3;\n\t}"
        : "+r"(aL), "+r"(aG), "+r"(sL), "+r"(sG)
        : "r"(our), "r"(rf), "r"(rn), "r"(dx), "r"(qb2), "r"(wa), "r"(ws), "r"(one));
}
// The undo half alone, for the steps that have no add part (rank-level undo steps): valid as above;
//   if (our < rn) sL += ws;  if (our > rn) sG += ws.
__device__ __forceinline__ void undo_step(uint32_t &sL, uint32_t &sG, uint32_t our, uint32_t rn, uint32_t dx, uint32_t qb2,
                                          uint32_t ws, uint32_t one) {
    asm("{\n\t.reg .pred v, p;\n\t.reg .b32 t;\n\t"
        "and.b32 t, %4, %5;\n\t"
        "setp.eq.u32 v, t, 0;\n\t"
        "setp.lt.and.u32 p, %2, %3, v;\n\t@p mad.lo.u32 %0, %6, %7, %0;\n\t"
        "setp.gt.and.u32 p, %2, %3, v;\n\t@p mad.lo.u32 %1, %6, %7, %1;\n\t}"
        : "+r"(sL), "+r"(sG)
        : "r"(our), "r"(rn), "r"(dx), "r"(qb2), "r"(ws), "r"(one));
}
// if (a < b) accL += w * m;  if (a > b) accG += w * m   (same pipe split)
__device__ __forceinline__ void cmp_add(uint32_t &accL, uint32_t &accG, uint32_t a, uint32_t b, uint32_t w, uint32_t one) {
    asm("{\n\t.reg .pred p;\n\t"
        "setp.lt.u32 p, %2, %3;\n\t@p mad.lo.u32 %0, %4, %5, %0;\n\t"
        "setp.gt.u32 p, %2, %3;\n\t@p mad.lo.u32 %1, %4, %5, %1;\n\t}"
        : "+r"(accL), "+r"(accG)
        : "r"(a), "r"(b), "r"(w), "r"(one));
}
// cmp_add that also records the two outcomes as bit `bit` of lb / gb (the compare bits feed the bit-parallel undo half)
__device__ __forceinline__ void cmp_add_bits(uint32_t &accL, uint32_t &accG, uint32_t &lb, uint32_t &gb, uint32_t a, uint32_t b,
                                             uint32_t w, uint32_t one, uint32_t bit) {
    asm("{\n\t.reg .pred p;\n\t"
        "setp.lt.u32 p, %4, %5;\n\t@p mad.lo.u32 %0, %6, %7, %0;\n\t@p or.b32 %2, %2, %8;\n\t"
        "setp.gt.u32 p, %4, %5;\n\t@p mad.lo.u32 %1, %6, %7, %1;\n\t@p or.b32 %3, %3, %8;\n\t}"
        : "+r"(accL), "+r"(accG), "+r"(lb), "+r"(gb)
        : "r"(a), "r"(b), "r"(w), "r"(one), "r"(bit));
}

}  // namespace ms
}  // namespace holdem
