// Pieces shared by the two symmetric half-band SpMM translation units (spmm_dsym.cu: one tile per CTA + finish
// kernel; spmm_dsymp.cu: persistent row ranges, single kernel): the register-level forward / transposed FMA
// groups and the per-warp bulk-copy ring primitives.
#pragma once
#include "sgv_device.cuh"

#define DS_FWD(C, XA, XB, XC, XD)                                                             \
    do {                                                                                      \
        const double v0 = (double)(C).x, v1 = (double)(C).y, v2 = (double)(C).z, v3 = (double)(C).w; \
        acc0.x = fma(v0, (XA).x, acc0.x); acc0.y = fma(v0, (XA).y, acc0.y);                   \
        acc1.x = fma(v1, (XB).x, acc1.x); acc1.y = fma(v1, (XB).y, acc1.y);                   \
        acc2.x = fma(v2, (XC).x, acc2.x); acc2.y = fma(v2, (XC).y, acc2.y);                   \
        acc3.x = fma(v3, (XD).x, acc3.x); acc3.y = fma(v3, (XD).y, acc3.y);                   \
    } while (0)

// transposed use of the 4 values of one diagonal: element e of diagonal d+k goes to target e+k
#define DS_TRN(C, TA, TB, TC, TD)                                                             \
    do {                                                                                      \
        const double v0 = (double)(C).x, v1 = (double)(C).y, v2 = (double)(C).z, v3 = (double)(C).w; \
        TA.x = fma(v0, O0.x, TA.x); TA.y = fma(v0, O0.y, TA.y);                               \
        TB.x = fma(v1, O1.x, TB.x); TB.y = fma(v1, O1.y, TB.y);                               \
        TC.x = fma(v2, O2.x, TC.x); TC.y = fma(v2, O2.y, TC.y);                               \
        TD.x = fma(v3, O3.x, TD.x); TD.y = fma(v3, O3.y, TD.y);                               \
    } while (0)

__device__ __forceinline__ double2 shfl_down1(double2 v) {   // lane 31 gets its own value back
    double2 r;
    r.x = __shfl_down_sync(0xffffffffu, v.x, 1);
    r.y = __shfl_down_sync(0xffffffffu, v.y, 1);
    return r;
}

static inline int ds_per(int Dp, int S) { return (((Dp + S - 1) / S) + 3) & ~3; }


// ---- per-warp TMA ring: the matrix stream goes global -> shared memory with bulk asynchronous copies
// (cp.async.bulk, completion on an mbarrier), so the bytes in flight cost no registers.  One stage = one
// group of 4 diagonals x the warp's 128 rows = 4 x 512 contiguous bytes.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

#define DS_STAGE_FLOATS 512   // 4 diagonals x 128 rows
#ifndef DS_O_IN_REGS
#define DS_O_IN_REGS 1
#endif


// The matrix is read exactly once per pass: mark its lines evict-first in L2 so that the stream does not push the
// (re-used) vectors out of the 126 MB L2.
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(unsigned dst, const void* src, unsigned bytes, unsigned bar, unsigned long long pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar), "l"(pol)
                 : "memory");
}
// Orders this thread's earlier generic-proxy accesses of shared memory (the LDS reads of a ring stage, made visible
// to it by __syncwarp) before its later async-proxy operations (the bulk copy that refills the stage).
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// pull `bytes` (multiple of 16) starting at a 16-byte aligned global address into L2 without a destination
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
