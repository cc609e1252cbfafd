// Symmetric half-band SpMM ("DSYM" layout): out = gamw * (R v) + gam2 * v for a pair of fp64
// vectors, R symmetric banded and stored ONCE: only the diagonals d = 0..w of the upper triangle,
// R[i][i+d] at storage row j = i + E (E: leading extension rows, see below), padded with zero diagonals to
// Dp = roundup(w+1,4) and zero rows to ldb = roundup(rows,128); diagonal 0 holds HALF of R[i][i] (exact in
// fp32), because the kernel uses every stored value twice - forward and transposed - and for d = 0 both
// uses hit row i.  The values are TILED so that what one warp consumes is one contiguous stream:
//     U[ ((j/128)*(Dp/4) + d/4)*512 + (d%4)*128 + j%128 ]        (sgv_dsym_index)
// i.e. per block of 128 rows, groups of 4 diagonals of 2 KB each, in diagonal order.  One pass reads 4*(w+1) bytes per row instead of
// the 4*(2w+1) of the full band - the matrix stream is the whole cost of a CG iteration (HBM bound),
// so this halves it.
//
// Every loaded value is used twice, from registers:
//   forward     y[i]   += U[d][i] * x[i+d]      x from a sliding register window over the shared-
//                                                memory x window (as in the full-band kernel)
//   transposed  y[i+d] += U[d][i] * x[i]        x[i] fixed in registers; the TARGET row moves with d.
// The transposed sums run as a systolic pipeline over the lanes of a warp.  A thread owns rows
// 4g..4g+3 and, for a group of 4 diagonals d..d+3, the 7 targets 4g+d..4g+d+6 in registers T0..T6.
// After the group T0..T3 are final for this thread; they are exactly the targets that the thread
// one lane below (rows 4(g-1)..) works on in ITS next group (4(g-1)+(d+4)+{0..3}), so they are
// handed down with one shuffle and seed that thread's accumulators; T4..T6 stay in the thread.
// Lane 0 emits 4 finished sums per group into a per-warp staging range, and after the last group
// every lane holds 4 finished sums for disjoint targets.  No atomics, no shared-memory
// read-modify-write in the loop, fixed summation order.
//
// A CTA (RW row-warps x S diagonal segments) then adds, in fixed order, the forward sums of its
// S segments and the staging ranges of its warps: rows of its own tile go to ypart[], the
// contributions to the Dp rows after the tile go to tails[cta][]; k_dsym_finish adds the (at most
// ceil(Dp/TR)) tails that reach a row, applies the fused epilogue and the grid / cross-rank reduction.
// (Consuming the tails inside the main kernel - tiles ordered by an atomic ticket, per-tile flags - was
// measured: the latency-bound tail of every CTA cost 0.43 ms per pass against 0.335 + 0.025 ms.)
//
// CG mode is a whole conjugate-gradient step in these two kernels (scipy.sparse.linalg.cg semantics, 2 RHS):
// while the x window is staged, the pending update r = r - alpha q is applied and the new direction
// p = r + beta p formed on the fly (also for the halo entries, from the neighbours' r, q, p), the owned
// rows of r, p and x += alpha p are written; the finish kernel writes q = A p and sums p.q, r.q, q.q, r.r, from
// which the finaliser gets alpha and |r - alpha q|^2 = r.r - 2 alpha r.q + alpha^2 q.q for the stopping test
// and beta - so a CG iteration costs one matrix pass, one small kernel and ONE reduction (no vector-update
// kernel, no second reduction).  r, p, q are double-buffered
// (step n reads buffers (n+1)&1, writes n&1), so no rank overwrites what a neighbour still reads.
//
// Row partition over GPUs: a rank also stores the E = roundup(w,256) rows BEFORE its first own row
// (entries that couple them to its own rows; the rest zero) and computes their transposed
// contributions itself, so no partial sums cross ranks - only vector halos are read from the
// neighbours' memory when the x window is staged (left halo for the extension rows, right halo for
// the forward part).  Results for the extension rows themselves are discarded.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include "sgv_device.cuh"

#include "dsym_common.cuh"

size_t sgv_dsym_smem_bytes(int64_t w, int rw, int s, int nst) {
    const int Dp = (int)round_up(w + 1, 4);
    const int TR = 128 * rw;
    size_t b = (size_t)4 * dia_plane_len(TR + Dp) * sizeof(double2);                   // x window
    b += (size_t)rw * s * (ds_per(Dp, s) + 128) * sizeof(double2);                      // per-warp staging of the transposed sums
    const size_t ring = (size_t)rw * s * nst * DS_STAGE_FLOATS * sizeof(float);         // per-warp TMA ring ...
    const size_t fwd = (size_t)s * TR * sizeof(double2);                                // ... reused for the forward sums
    b += ring > fwd ? ring : fwd;
    b += (size_t)rw * s * nst * 8;                                                      // mbarriers
    return b;
}

template <int RW, int S, int NST, int MINB, bool CG>
__global__ void __launch_bounds__(32 * RW * S, MINB)
k_spmm_dsym(SpmmArgs a, const float* __restrict__ U, int Dp, int64_t ldb, int64_t E, double2* __restrict__ ypart,
            double2* __restrict__ tails) {
    if (a.check_done && a.rc.st->done[0] && a.rc.st->done[1]) return;
    // Tiles whose x window reaches into the right neighbour's rows stage their window over NVLink (a few
    // microseconds of latency instead of an L2 hit): they go FIRST, so that this overlaps the streaming of the
    // other tiles instead of stretching the tail of the kernel.  (The left-edge tiles are the first ones anyway.)
    int tile = blockIdx.x;
    if (a.v_right != nullptr) {
        const int ntiles = gridDim.x;
        const int nedge = min(ntiles, (Dp + 128 * RW - 1) / (128 * RW) + 1);
        tile = (int)blockIdx.x < nedge ? ntiles - 1 - (int)blockIdx.x : (int)blockIdx.x - nedge;
    }
    constexpr int TR = 128 * RW;
    constexpr int NT = 32 * RW * S;
    constexpr int NW = RW * S;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int W = TR + Dp;
    const int PL = dia_plane_len(W);
    const int per = (((Dp + S - 1) / S) + 3) & ~3;
    const int SL = per + 128;
    constexpr size_t RING_B = (size_t)NW * NST * DS_STAGE_FLOATS * sizeof(float);
    constexpr size_t FWD_B = (size_t)S * TR * sizeof(double2);
    float* ring = reinterpret_cast<float*>(smem_raw);                                   // 128-byte aligned stages
    double2* fwd = reinterpret_cast<double2*>(smem_raw);                                // aliases the ring (used after the loop)
    double2* xw = reinterpret_cast<double2*>(smem_raw + (RING_B > FWD_B ? RING_B : FWD_B));
    double2* stag = xw + 4 * PL;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(stag + NW * SL);

    const int64_t r0s = (int64_t)tile * TR;         // storage index of the tile's first row
    const int64_t r0 = r0s - E;                     // the same in local coordinates (0 = first own row)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int rw = wid % RW, s = wid / RW;
    const int g = rw * 32 + lane;
    const int d0 = s * per;
    const int d1 = min(Dp, d0 + per);
    const int64_t wrow = r0s + 128 * rw;            // first row of this warp (ldb is a multiple of 128)
    // warp-uniform; Dp and per are multiples of 4.  A warp beyond the stored rows has no work.
    const int ngroups = (d0 < d1 && wrow < ldb) ? ((d1 - d0) >> 2) : 0;

    // arm the ring: lane 0 initialises the warp's barriers and issues one 2 KB bulk copy per stage
    float* wring = ring + (size_t)wid * NST * DS_STAGE_FLOATS;
    const unsigned bar0 = smem_u32(bars + wid * NST);
    const unsigned ring0 = smem_u32(wring);
    const float* gsrc = U + ((wrow >> 7) * (int64_t)(Dp >> 2) + (d0 >> 2)) * DS_STAGE_FLOATS;   // group gi: + gi*512 floats
    if (ngroups > 0 && lane == 0) {
#pragma unroll
        for (int t = 0; t < NST; ++t) mbar_init(bar0 + 8 * t, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
        for (int t = 0; t < NST; ++t) {
            if (t < ngroups) {
                mbar_expect_tx(bar0 + 8 * t, DS_STAGE_FLOATS * 4);
                bulk_g2s(ring0 + t * DS_STAGE_FLOATS * 4, gsrc + (int64_t)t * DS_STAGE_FLOATS, DS_STAGE_FLOATS * 4, bar0 + 8 * t);
            }
        }
    }
    __syncwarp();   // the barriers are initialised before any lane waits on them

    // stage the x window [r0, r0+TR+Dp) (local coordinates).  Entries left of 0 come from the left
    // neighbour (extension rows), entries right of M from the right neighbour, zero at the matrix edges.
    // CG mode: the window holds the new CG direction, computed on the fly (see the header comment).
    double al0 = 0.0, al1 = 0.0, beta0 = 0.0, beta1 = 0.0;
    bool first = true, fz0 = false, fz1 = false;
    if (CG) {
        const CgState* st = a.rc.st;
        first = st->step == 0;
        fz0 = st->done[0] != 0;
        fz1 = st->done[1] != 0;
        if (!first) {
            al0 = st->alpha[0];
            al1 = st->alpha[1];
            beta0 = st->rho[0] / st->rho_prev[0];
            beta1 = st->rho[1] / st->rho_prev[1];
        }
    }
    for (int j = threadIdx.x; j < 4 * PL; j += NT) {
        const int64_t col = r0 + j;
        double2 val = make_double2(0.0, 0.0);
        if (j < W) {
            const double2 *src = nullptr, *rsrc = nullptr, *qsrc = nullptr;
            int64_t idx = col;
            if (col >= 0 && col < a.M) {
                src = a.v;
                rsrc = a.r;
                qsrc = a.q;
            } else if (col < 0 && a.v_left != nullptr && a.n_left + col >= 0) {
                src = a.v_left;
                rsrc = a.r_left;
                qsrc = a.q_left;
                idx = a.n_left + col;
            } else if (col >= a.M && a.v_right != nullptr) {
                src = a.v_right;
                rsrc = a.r_right;
                qsrc = a.q_right;
                idx = col - a.M;
            }
            if (src != nullptr) {
                if (!CG) {
                    val = ld_vec2(src + idx);
                } else {
                    double2 rv = ld_vec2(rsrc + idx);
                    double2 po = make_double2(0.0, 0.0);
                    if (!first) {                       // nothing is pending before the first step (p, q not yet defined)
                        po = ld_vec2(src + idx);
                        const double2 qo = ld_vec2(qsrc + idx);
                        if (al0 != 0.0) rv.x -= al0 * qo.x;     // scipy: r -= alpha*q
                        if (al1 != 0.0) rv.y -= al1 * qo.y;
                    }
                    // scipy: p *= beta; p += r   (first step: p = r); a finished column keeps its p
                    val.x = fz0 ? po.x : (first ? rv.x : po.x * beta0 + rv.x);
                    val.y = fz1 ? po.y : (first ? rv.y : po.y * beta1 + rv.y);
                    if (j < TR && col >= 0 && col < a.M) {      // owned rows of this tile
                        a.r_new[col] = rv;
                        a.p_new[col] = val;
                        if (al0 != 0.0 || al1 != 0.0) {
                            double2 xv = a.x[col];
                            if (al0 != 0.0) xv.x += al0 * po.x; // scipy: x += alpha*p
                            if (al1 != 0.0) xv.y += al1 * po.y;
                            a.x[col] = xv;
                        }
                    }
                }
            }
        }
        xw[(j & 3) * PL + (j >> 2)] = val;
    }
    __syncthreads();

    double2 acc0 = make_double2(0, 0), acc1 = acc0, acc2 = acc0, acc3 = acc0;
    if (ngroups > 0) {
        double2 T0 = acc0, T1 = acc0, T2 = acc0, T3 = acc0, T4 = acc0, T5 = acc0, T6 = acc0;
#if DS_O_IN_REGS
        const double2 O0 = xw[g], O1 = xw[PL + g], O2 = xw[2 * PL + g], O3 = xw[3 * PL + g];   // x[row4 .. row4+3]
#endif
        int xi = g + (d0 >> 2);
        double2 X0 = xw[xi], X1 = xw[PL + xi], X2 = xw[2 * PL + xi], X3 = xw[3 * PL + xi];
        double2* st = stag + wid * SL;
        const double hm = lane == 31 ? 0.0 : 1.0;   // lane 31 has no lane above: its incoming sums are zero
        for (int gb = 0; gb < ngroups; gb += NST) {
            const unsigned parity = (unsigned)(gb / NST) & 1u;
#pragma unroll
            for (int stage = 0; stage < NST; ++stage) {
                const int gi = gb + stage;
                if (gi < ngroups) {
                    // take this group's 4 x float4 out of the ring, then hand the stage straight back to the copy engine
                    mbar_wait(bar0 + 8 * stage, parity);
                    const float4* sp = reinterpret_cast<const float4*>(wring + stage * DS_STAGE_FLOATS) + lane;
                    const float4 c0 = sp[0], c1 = sp[32], c2 = sp[64], c3 = sp[96];
                    __syncwarp();
                    if (lane == 0 && gi + NST < ngroups) {
                        mbar_expect_tx(bar0 + 8 * stage, DS_STAGE_FLOATS * 4);
                        bulk_g2s(ring0 + stage * DS_STAGE_FLOATS * 4, gsrc + (int64_t)(gi + NST) * DS_STAGE_FLOATS,
                                 DS_STAGE_FLOATS * 4, bar0 + 8 * stage);
                    }
                    const double2 N0 = xw[xi + 1], N1 = xw[PL + xi + 1], N2 = xw[2 * PL + xi + 1], N3 = xw[3 * PL + xi + 1];
#if !DS_O_IN_REGS
                    // x[row4 .. row4+3] re-read per group: 4 conflict-free LDS.128 instead of 16 live registers
                    const double2 O0 = xw[g], O1 = xw[PL + g], O2 = xw[2 * PL + g], O3 = xw[3 * PL + g];
#endif
                    DS_FWD(c0, X0, X1, X2, X3);
                    DS_TRN(c0, T0, T1, T2, T3);
                    DS_FWD(c1, X1, X2, X3, N0);
                    DS_TRN(c1, T1, T2, T3, T4);
                    DS_FWD(c2, X2, X3, N0, N1);
                    DS_TRN(c2, T2, T3, T4, T5);
                    DS_FWD(c3, X3, N0, N1, N2);
                    DS_TRN(c3, T3, T4, T5, T6);
                    X0 = N0; X1 = N1; X2 = N2; X3 = N3;
                    ++xi;
                    // hand the 4 finished sums down one lane; lane 0's are final for the warp
                    if (lane == 0) {
                        double2* e = st + 4 * gi;
                        e[0] = T0; e[1] = T1; e[2] = T2; e[3] = T3;
                    }
                    const double2 I0 = shfl_down1(T0), I1 = shfl_down1(T1), I2 = shfl_down1(T2), I3 = shfl_down1(T3);
                    T0 = make_double2(fma(I0.x, hm, T4.x), fma(I0.y, hm, T4.y));
                    T1 = make_double2(fma(I1.x, hm, T5.x), fma(I1.y, hm, T5.y));
                    T2 = make_double2(fma(I2.x, hm, T6.x), fma(I2.y, hm, T6.y));
                    T3 = make_double2(I3.x * hm, I3.y * hm);
                    T4 = T5 = T6 = make_double2(0.0, 0.0);
                }
            }
        }
        // drain: every lane now holds finished sums for the disjoint targets (d1-d0) + 4*lane + {0..3}
        double2* e = st + (d1 - d0) + 4 * lane;
        e[0] = T0; e[1] = T1; e[2] = T2; e[3] = T3;
    }
    __syncthreads();   // every warp is done with its ring: the forward sums reuse that memory
    {
        double2* f = fwd + s * TR + 4 * g;
        f[0] = acc0; f[1] = acc1; f[2] = acc2; f[3] = acc3;
    }
    __syncthreads();

    // fixed-order combination: forward sums of the S segments + the staging ranges that cover the target
    double2* mytails = tails + (int64_t)tile * Dp;
    for (int t = threadIdx.x; t < TR + Dp; t += NT) {
        double2 sum = make_double2(0.0, 0.0);
        if (t < TR) {
#pragma unroll
            for (int s2 = 0; s2 < S; ++s2) {
                const double2 v = fwd[s2 * TR + t];
                sum.x += v.x;
                sum.y += v.y;
            }
        }
#pragma unroll
        for (int s2 = 0; s2 < S; ++s2) {
            const int e0 = s2 * per, e1 = min(Dp, e0 + per);
            if (e0 < e1) {
#pragma unroll
                for (int rw2 = 0; rw2 < RW; ++rw2) {
                    const int rel = t - 128 * rw2 - e0;
                    if (rel >= 0 && rel < (e1 - e0) + 128 && r0s + 128 * rw2 < ldb) {
                        const double2 v = stag[(s2 * RW + rw2) * SL + rel];
                        sum.x += v.x;
                        sum.y += v.y;
                    }
                }
            }
        }
        if (t < TR) {
            if (r0s + t < ldb) ypart[r0s + t] = sum;
        } else {
            mytails[t - TR] = sum;
        }
    }
}

// y[i] = ypart[i] + the tails of the preceding tiles that reach row i; fused epilogue + reduction.
// EPI_CG: q = A p and the four dot products of the fused CG step.
template <int EPI>
__global__ void __launch_bounds__(256)
k_dsym_finish(SpmmArgs a, const double2* __restrict__ ypart, const double2* __restrict__ tails, int Dp, int TR, int64_t E,
              const double2* __restrict__ vin) {
    SGV_LOAD_DEV_SCALARS(a);
    if (a.check_done && a.rc.st->done[0] && a.rc.st->done[1]) return;
    constexpr int NV = EPI == EPI_CG ? 8 : 2;
    __shared__ double red[NV * 32];
    double dots[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) dots[k] = 0.0;
    // grid-stride: a few hundred blocks (one ticket atomic and one partial row each), not one per 256 rows
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.M; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t J = i + E;
        const int64_t b = J / TR;
        double2 y = ypart[J];
        for (int64_t bb = b - 1; bb >= 0; --bb) {
            const int64_t off = J - (bb + 1) * TR;
            if (off >= Dp) break;
            const double2 t = tails[bb * Dp + off];
            y.x += t.x;
            y.y += t.y;
        }
        const double2 vi = vin[i];
        if constexpr (EPI == EPI_CG) {
            double2 o;
            o.x = a.gamw * y.x + a.gam2 * vi.x;
            o.y = a.gamw * y.y + a.gam2 * vi.y;
            a.out[i] = o;                                    // q = A p
            const double2 rv = a.r_new[i];
            dots[0] += vi.x * o.x; dots[1] += vi.y * o.y;    // p.q
            dots[2] += rv.x * o.x; dots[3] += rv.y * o.y;    // r.q
            dots[4] += o.x * o.x;  dots[5] += o.y * o.y;     // q.q
            dots[6] += rv.x * rv.x; dots[7] += rv.y * rv.y;  // r.r
        } else {
            epi_row<EPI>(a, i, y, vi, dots);
        }
    }
    if constexpr (EPI != EPI_PLAIN) grid_reduce<NV>(dots, a.rc, red);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
#define DS_BIG_RW 2
#define DS_BIG_S 4
#define DS_NST 4       // ring stages per warp (2 KB each): 16 warps x 8 KB = 128 KB of matrix bytes in flight per SM
#define DS_MINB 2
#define DS_SMEM_LIMIT (220 * 1024)

bool sgv_dsym_feasible(int64_t w) { return sgv_dsym_smem_bytes(w, 1, 8, DS_NST) <= DS_SMEM_LIMIT; }

static bool ds_use_big(const sgv_ctx* c, const LdMatrix& ld) {
    static const long long waves = getenv("SGV_DS_BIG_WAVES") ? atoll(getenv("SGV_DS_BIG_WAVES")) : 1;   // tuning knob
    return ld.ldb >= (int64_t)c->sm_count * 2 * 256 * waves && sgv_dsym_smem_bytes(ld.w, DS_BIG_RW, DS_BIG_S, DS_NST) <= 112 * 1024;
}

template <bool CG>
static int preload_main() {
    SGV_CUDA(cudaFuncSetAttribute(k_spmm_dsym<DS_BIG_RW, DS_BIG_S, DS_NST, DS_MINB, CG>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, DS_SMEM_LIMIT));
    SGV_CUDA(cudaFuncSetAttribute(k_spmm_dsym<1, 8, DS_NST, DS_MINB, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  DS_SMEM_LIMIT));
    return 0;
}

int sgv_preload_dsym() {
    SGV_TRY(preload_main<false>());
    SGV_TRY(preload_main<true>());
    SGV_TRY(sgv_preload_dsymp());
    cudaFuncAttributes fa;
    SGV_CUDA(cudaFuncGetAttributes(&fa, k_dsym_finish<EPI_Q>));
    SGV_CUDA(cudaFuncGetAttributes(&fa, k_dsym_finish<EPI_RESID>));
    SGV_CUDA(cudaFuncGetAttributes(&fa, k_dsym_finish<EPI_STATS>));
    SGV_CUDA(cudaFuncGetAttributes(&fa, k_dsym_finish<EPI_PLAIN>));
    SGV_CUDA(cudaFuncGetAttributes(&fa, k_dsym_finish<EPI_CG>));
    return 0;
}

// scratch (per handle, sized for the largest DSYM matrix uploaded so far); called at upload / adopt time
// only, never inside the solver loop
int sgv_dsym_ensure_scratch(sgv_ctx* c, const LdMatrix& ld) {
    const int64_t Dp = round_up(ld.w + 1, 4);
    const int64_t tiles = (ld.ldb + 127) / 128;   // upper bound for either tile shape
    const int64_t need_t = tiles * Dp;
    if (c->ds_tails_cap < need_t) {
        SGV_CUDA(cudaStreamSynchronize(c->stream));
        if (c->ds_tails) cudaFree(c->ds_tails);
        c->ds_tails = nullptr;
        c->ds_tails_cap = 0;
        SGV_CUDA(cudaMalloc(&c->ds_tails, need_t * sizeof(double2)));
        c->ds_tails_cap = need_t;
    }
    if (c->ds_ypart_cap < ld.ldb) {
        SGV_CUDA(cudaStreamSynchronize(c->stream));
        if (c->ds_ypart) cudaFree(c->ds_ypart);
        c->ds_ypart = nullptr;
        c->ds_ypart_cap = 0;
        SGV_CUDA(cudaMalloc(&c->ds_ypart, ld.ldb * sizeof(double2)));
        c->ds_ypart_cap = ld.ldb;
    }
    return sgv_dsymp_ensure_scratch(c, ld);
}

template <int RW, int S, bool CG>
static int launch_main(sgv_ctx* c, const LdMatrix& ld, const SpmmArgs& a) {
    constexpr int TR = 128 * RW;
    const size_t smem = sgv_dsym_smem_bytes(ld.w, RW, S, DS_NST);
    SGV_CHECK(smem <= DS_SMEM_LIMIT, "half-bandwidth %lld too large for the DSYM kernel", (long long)ld.w);
    const unsigned grid = (unsigned)((ld.ldb + TR - 1) / TR);
    const int Dp = (int)round_up(ld.w + 1, 4);
    k_spmm_dsym<RW, S, DS_NST, DS_MINB, CG><<<grid, 32 * RW * S, smem, c->stream>>>(a, ld.band, Dp, ld.ldb, ld.ext, c->ds_ypart,
                                                                                   c->ds_tails);
    c->launches++;
    return 0;
}

template <int EPI>
static int launch_finish(sgv_ctx* c, const LdMatrix& ld, SpmmArgs& a, int TR, const double2* vin) {
    const unsigned grid = (unsigned)std::min<int64_t>((c->Ml + 255) / 256, (int64_t)c->sm_count * 6);
    SGV_TRY(sgv_ensure_partials(c, grid));
    a.rc.partials = c->partials;
    const int Dp = (int)round_up(ld.w + 1, 4);
    k_dsym_finish<EPI><<<grid, 256, 0, c->stream>>>(a, c->ds_ypart, c->ds_tails, Dp, TR, ld.ext, vin);
    c->launches++;
    return 0;
}

// a: fully prepared SpmmArgs (vectors, halos, epilogue operands, reduction context)
int sgv_launch_dsym(sgv_ctx* c, const LdMatrix& ld, int epi, SpmmArgs& a) {
    SGV_CHECK(c->ds_ypart != nullptr && c->ds_tails != nullptr, "DSYM scratch not allocated");
    if ((epi == EPI_CG || epi == EPI_Q || epi == EPI_PLAIN) && sgv_dsymp_feasible(ld.w)) return sgv_launch_dsymp(c, ld, epi, a);
    const bool big = ds_use_big(c, ld);
    const int TR = big ? 128 * DS_BIG_RW : 128;
    if (epi == EPI_CG) {
        if (big) SGV_TRY((launch_main<DS_BIG_RW, DS_BIG_S, true>(c, ld, a)));
        else SGV_TRY((launch_main<1, 8, true>(c, ld, a)));
        return launch_finish<EPI_CG>(c, ld, a, TR, a.p_new);
    }
    if (big) SGV_TRY((launch_main<DS_BIG_RW, DS_BIG_S, false>(c, ld, a)));
    else SGV_TRY((launch_main<1, 8, false>(c, ld, a)));
    switch (epi) {
        case EPI_Q: return launch_finish<EPI_Q>(c, ld, a, TR, a.v);
        case EPI_RESID: return launch_finish<EPI_RESID>(c, ld, a, TR, a.v);
        case EPI_STATS: return launch_finish<EPI_STATS>(c, ld, a, TR, a.v);
        default: return launch_finish<EPI_PLAIN>(c, ld, a, TR, a.v);
    }
}

// One fused CG step (EPI_CG): step n reads r, p, q from buffers (n+1)&1 - the neighbours' too - and writes n&1.
int sgv_launch_dsym_cg(sgv_ctx* c, Cohort& co, int n, double gamw, double gam2) {
    const LdMatrix& ld = co.ld;
    SGV_CHECK(ld.layout == SGV_LAYOUT_DSYM, "fused CG step needs the DSYM layout");
    const int prev = (n + 1) & 1, cur = n & 1;
    SpmmArgs a;
    memset(&a, 0, sizeof(a));
    a.vs_cohort = c->vs_active;   // >= 0 inside the fused VAMP iteration: gamw / gam2 come from the device
    a.v = c->pp[prev];
    a.r = c->rr2[prev];
    a.q = c->qq2[prev];
    a.p_new = c->pp[cur];
    a.r_new = c->rr2[cur];
    a.out = c->qq2[cur];
    a.x = c->xx;
    if (c->world > 1 && c->halo) {
        if (c->rank > 0) {
            const PeerView& pv = c->peer[c->rank - 1];
            SGV_CHECK(pv.base != nullptr && pv.Ml >= ld.w, "left neighbour not attached or shorter than the half-bandwidth");
            a.v_left = reinterpret_cast<double2*>(pv.base + arena_off_pp(pv.Ml, prev));
            a.r_left = reinterpret_cast<double2*>(pv.base + arena_off_rr(pv.Ml, prev));
            a.q_left = reinterpret_cast<double2*>(pv.base + arena_off_qq(pv.Ml, prev));
            a.n_left = pv.Ml;
        }
        if (c->rank + 1 < c->world) {
            const PeerView& pv = c->peer[c->rank + 1];
            SGV_CHECK(pv.base != nullptr && pv.Ml >= ld.w, "right neighbour not attached or shorter than the half-bandwidth");
            a.v_right = reinterpret_cast<double2*>(pv.base + arena_off_pp(pv.Ml, prev));
            a.r_right = reinterpret_cast<double2*>(pv.base + arena_off_rr(pv.Ml, prev));
            a.q_right = reinterpret_cast<double2*>(pv.base + arena_off_qq(pv.Ml, prev));
        }
    }
    a.bb = c->bb;
    a.gamw = gamw;
    a.gam2 = gam2;
    a.M = c->Ml;
    a.check_done = 1;
    a.rc = sgv_red_begin(c, AP_CGFUSED, 8, 0);
    a.rc.skip_if_done = SKIP_CG_DONE;
    a.rc.st = c->cg;
    if (c->prof) {
        if (c->prof_n + 2 > c->prof_ev.size()) {
            for (int i = 0; i < 256; ++i) {
                cudaEvent_t e;
                SGV_CUDA(cudaEventCreate(&e));
                c->prof_ev.push_back(e);
            }
        }
        SGV_CUDA(cudaEventRecord(c->prof_ev[c->prof_n], c->stream));
    }
    SGV_TRY(sgv_launch_dsym(c, ld, EPI_CG, a));
    if (c->prof) {
        SGV_CUDA(cudaEventRecord(c->prof_ev[c->prof_n + 1], c->stream));
        c->prof_n += 2;
    }
    SGV_CUDA(cudaGetLastError());
    return sgv_red_end(c, a.rc);
}

// The whole CG solve as ONE cooperative launch of the persistent kernel (spmm_dsymp.cu): no launch per step, no host
// read-back inside the solve.  The cross-rank exchanges of its steps are numbered on the device (RedCtx::pubseq).
int sgv_launch_dsym_solve(sgv_ctx* c, Cohort& co, double gamw, double gam2, int max_steps) {
    const LdMatrix& ld = co.ld;
    SGV_CHECK(sgv_dsymp_solve_usable(c, ld), "whole-solve kernel not usable for this cohort");
    SpmmArgs a;
    memset(&a, 0, sizeof(a));
    a.vs_cohort = c->vs_active;   // >= 0 inside the fused VAMP iteration: gamw / gam2 come from the device
    a.x = c->xx;
    a.bb = c->bb;
    a.gamw = gamw;
    a.gam2 = gam2;
    a.M = c->Ml;
    a.rc = sgv_red_begin(c, AP_CGFUSED, 8, 0);
    a.rc.st = c->cg;
    if (c->prof) {
        if (c->prof_n + 2 > c->prof_ev.size()) {
            for (int i = 0; i < 256; ++i) {
                cudaEvent_t e;
                SGV_CUDA(cudaEventCreate(&e));
                c->prof_ev.push_back(e);
            }
        }
        SGV_CUDA(cudaEventRecord(c->prof_ev[c->prof_n], c->stream));
    }
    SGV_TRY(sgv_launch_dsymp_solve(c, ld, a, max_steps));
    if (c->prof) {
        SGV_CUDA(cudaEventRecord(c->prof_ev[c->prof_n + 1], c->stream));
        c->prof_n += 2;
    }
    SGV_CUDA(cudaGetLastError());
    return 0;
}
