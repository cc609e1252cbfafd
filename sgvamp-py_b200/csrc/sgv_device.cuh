// Device-side helpers: streaming loads, deterministic block / grid reductions.
#pragma once
#include "sgv_internal.cuh"

#define SGV_MAX_PARTIAL_VALUES 16

__device__ __forceinline__ float4 ldg_stream_f4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_stream_f1(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ int ldg_stream_i1(const int* p) {
    int r;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Sum NV values over the block in a fixed order.  Result valid in thread 0 (all lanes of warp 0).
// `red` is shared scratch of at least NV*32 doubles.  Contains __syncthreads().
template <int NV, bool MIN = false>
__device__ __forceinline__ void block_reduce(double (&v)[NV], double* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = MIN ? warp_min(v[k]) : warp_sum(v[k]);
    __syncthreads();   // protect `red` from a previous use
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) red[k * 32 + wid] = v[k];
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double x = (lane < nw) ? red[k * 32 + lane] : (MIN ? __longlong_as_double(0x7ff0000000000000LL) : 0.0);
            v[k] = MIN ? warp_min(x) : warp_sum(x);
        }
    }
}

// Grid-wide deterministic reduction: every block contributes NV values; the block that takes the
// last ticket sums the per-block partials in index order and calls fin(totals) from thread 0.
// Works for any grid shape; `partials` must hold NV * (number of blocks) doubles and *counter must
// be 0 on entry (it is reset to 0 on exit).
template <int NV, bool MIN = false, class Fin>
__device__ __forceinline__ void grid_reduce(double (&v)[NV], double* partials, unsigned* counter, double* red,
                                            Fin fin) {
    __shared__ int s_last;
    const unsigned nblk = gridDim.x * gridDim.y * gridDim.z;
    const unsigned bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    block_reduce<NV, MIN>(v, red);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) partials[(size_t)bid * NV + k] = v[k];
        __threadfence();
        unsigned t = atomicAdd(counter, 1u);
        s_last = (t == nblk - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = MIN ? __longlong_as_double(0x7ff0000000000000LL) : 0.0;
    for (unsigned b = threadIdx.x; b < nblk; b += blockDim.x) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double x = __ldcg(&partials[(size_t)b * NV + k]);
            acc[k] = MIN ? fmin(acc[k], x) : acc[k] + x;
        }
    }
    block_reduce<NV, MIN>(acc, red);
    if (threadIdx.x == 0) {
        fin(acc);
        *counter = 0u;
    }
}

// scipy.sparse.linalg.cg loop-top test for one column, applied right after rho = r.r is known.
// (scipy 1.18.1 _isolve/iterative.py: `if np.linalg.norm(r) < atol: return x, 0`, atol = 1e-5*|b|;
//  loop exhaustion returns info = maxiter without a final test.)
__device__ __forceinline__ void cg_top_test(CgState* s, int c) {
    if (s->done[c]) return;
    if (s->iters[c] >= s->maxit) {
        s->done[c] = 1;
        s->info[c] = s->maxit;
        return;
    }
    if (sqrt(s->rho[c]) < 1e-5 * sqrt(s->bnorm2[c])) {
        s->done[c] = 1;
        s->info[c] = 0;
    }
}
