// Fused shifted SpMM kernels:  out = gamw * (R v) + gam2 * v  for a pair of fp64 vectors (the
// xhat2 system and the Hutchinson probe system of the LMMSE step share the matrix, reference
// src/sgvamp.py:312-332), with the CG dot products / residual fused into the epilogue.
//
// R is stored fp32 in one of three layouts (DESIGN.md "Data layout in HBM"):
//   DIA        diagonal-major band, band[d*ldb + i] = R[i][i+d-w]        4 B / stored value
//   panels     row-major dense blocks (whole matrix, or one per LD block) 4 B / stored value
//   CSR        fp32 value + int32 column                                  8 B / stored value
// All three are HBM-bound streaming kernels (0.5 - 1 flop/B): 128-bit coalesced loads of the
// matrix, vector operands in shared memory / registers, fp64 FMA accumulation.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include "sgv_device.cuh"

// ---------------------------------------------------------------------------------------------
// DIA (banded) kernel
//   CTA = RW row-warps x S diagonal-segments; a thread owns 4 consecutive rows (one float4 per
//   diagonal) and a contiguous range of diagonals; the x window [r0-w, r0+TR+w) sits in shared
//   memory split into 4 planes by (index mod 4) so that the stride-4 access of consecutive lanes
//   is bank-conflict free; each thread slides a 4-entry register window along the diagonals, so a
//   step of 4 rows x 1 diagonal x 2 RHS costs 1 LDG.128 + 1 LDS.128 + 4 cvt + 8 DFMA.
// ---------------------------------------------------------------------------------------------
size_t sgv_dia_smem_bytes(int64_t w, int rw, int s) {
    int TR = 128 * rw;
    int W = TR + 2 * (int)w;
    size_t b = (size_t)4 * dia_plane_len(W) * sizeof(double2);
    b += (size_t)(s > 1 ? (s - 1) : 0) * 8 * (TR / 4) * sizeof(double);
    b += 2 * 32 * sizeof(double);
    return b;
}
bool sgv_dia_feasible(int64_t w) { return sgv_dia_smem_bytes(w, 1, 8) <= 200 * 1024; }

#define FMA4(C, XA, XB, XC, XD)                                       \
    do {                                                              \
        double v0 = (double)(C).x, v1 = (double)(C).y, v2 = (double)(C).z, v3 = (double)(C).w; \
        acc0.x = fma(v0, (XA).x, acc0.x); acc0.y = fma(v0, (XA).y, acc0.y); \
        acc1.x = fma(v1, (XB).x, acc1.x); acc1.y = fma(v1, (XB).y, acc1.y); \
        acc2.x = fma(v2, (XC).x, acc2.x); acc2.y = fma(v2, (XC).y, acc2.y); \
        acc3.x = fma(v3, (XD).x, acc3.x); acc3.y = fma(v3, (XD).y, acc3.y); \
    } while (0)

template <int RW, int S, int EPI, int PF, int MINB>
__global__ void __launch_bounds__(32 * RW * S, MINB)
k_spmm_dia(SpmmArgs a, const float* __restrict__ band, int w, int64_t ldb) {
    SGV_LOAD_DEV_SCALARS(a);
    if (a.check_done && a.rc.st->done[0] && a.rc.st->done[1]) return;
    constexpr int TR = 128 * RW;
    constexpr int NT = 32 * RW * S;
    extern __shared__ double2 smem2[];
    const int W = TR + 2 * w;
    const int PL = dia_plane_len(W);
    double2* xw = smem2;
    double* redseg = reinterpret_cast<double*>(xw + 4 * PL);
    double* red = redseg + (S > 1 ? (S - 1) : 0) * 8 * (TR / 4);

    const int64_t r0 = (int64_t)blockIdx.x * TR;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int rw = wid % RW, s = wid / RW;
    const int g = rw * 32 + lane;
    const int64_t row4 = r0 + 4 * g;
    const bool active = row4 < ldb;

    const int Dtot = 2 * w + 1;
    const int per = (((Dtot + S - 1) / S) + 3) & ~3;
    const int d0 = s * per;
    const int d1 = min(Dtot, d0 + per);

    // Issue the first PF groups of matrix loads before the x window is staged: the HBM latency of the
    // ring fill overlaps the staging pass and its barrier.
    const bool work = active && d0 < d1;
    const float* bp = band + (int64_t)d0 * ldb + row4;
    const int nfull = work ? ((d1 - d0) >> 2) : 0;
    // ring of PF groups (4 diagonals each) in flight: a slot is refilled right after it is consumed
    float4 q[PF][4];
#pragma unroll
    for (int j = 0; j < PF; ++j) {
        if (j < nfull) {
            const float* lp = bp + (int64_t)(4 * j) * ldb;
            q[j][0] = ldg_stream_f4(lp);
            q[j][1] = ldg_stream_f4(lp + ldb);
            q[j][2] = ldg_stream_f4(lp + 2 * ldb);
            q[j][3] = ldg_stream_f4(lp + 3 * ldb);
        }
    }

    // stage the x window [r0-w, r0+TR+w) in local coordinates; entries left of 0 / right of M come
    // from the neighbouring ranks' vectors through peer memory (or are zero at the matrix edge).
    // In fused mode the window holds the new CG direction p = r + beta*p_old, computed on the fly
    // (also for the halo entries, from the neighbours' r and p_old), and the owned part is written
    // to p_new - no separate direction-update kernel and no halo exchange step.
    double beta0 = 0.0, beta1 = 0.0;
    bool first = true, fz0 = false, fz1 = false;
    if (a.fused_p) {
        const CgState* st = a.rc.st;
        first = st->step == 0;
        fz0 = st->done[0] != 0;
        fz1 = st->done[1] != 0;
        if (!first) {
            beta0 = st->rho[0] / st->rho_prev[0];
            beta1 = st->rho[1] / st->rho_prev[1];
        }
    }
    for (int j = threadIdx.x; j < 4 * PL; j += NT) {
        const int64_t col = r0 - w + j;
        double2 val = make_double2(0.0, 0.0);
        if (j < W) {
            const double2 *src = nullptr, *rsrc = nullptr;
            int64_t idx = col;
            if (col >= 0 && col < a.M) {
                src = a.v;
                rsrc = a.r;
            } else if (col < 0 && a.v_left != nullptr) {
                src = a.v_left;
                rsrc = a.r_left;
                idx = a.n_left + col;
            } else if (col >= a.M && a.v_right != nullptr) {
                src = a.v_right;
                rsrc = a.r_right;
                idx = col - a.M;
            }
            if (src != nullptr) {
                if (!a.fused_p) {
                    val = ld_vec2(src + idx);
                } else {
                    const double2 rv = ld_vec2(rsrc + idx);
                    double2 po = make_double2(0.0, 0.0);
                    if (!first || fz0 || fz1) po = ld_vec2(src + idx);
                    // scipy: p *= beta; p += z   (first step: p = z)
                    val.x = fz0 ? po.x : (first ? rv.x : po.x * beta0 + rv.x);
                    val.y = fz1 ? po.y : (first ? rv.y : po.y * beta1 + rv.y);
                    if (j >= w && j < w + TR && col < a.M) a.p_new[col] = val;
                }
            }
        }
        xw[(j & 3) * PL + (j >> 2)] = val;
    }
    __syncthreads();

    double2 acc0 = make_double2(0, 0), acc1 = acc0, acc2 = acc0, acc3 = acc0;
    if (work) {
        int xi = g + (d0 >> 2);
        double2 X0 = xw[xi], X1 = xw[PL + xi], X2 = xw[2 * PL + xi], X3 = xw[3 * PL + xi];
        for (int m = 0; m < nfull; m += PF) {
#pragma unroll
            for (int j = 0; j < PF; ++j) {
                if (m + j < nfull) {
                    const double2 N0 = xw[xi + 1], N1 = xw[PL + xi + 1], N2 = xw[2 * PL + xi + 1], N3 = xw[3 * PL + xi + 1];
                    FMA4(q[j][0], X0, X1, X2, X3);
                    FMA4(q[j][1], X1, X2, X3, N0);
                    FMA4(q[j][2], X2, X3, N0, N1);
                    FMA4(q[j][3], X3, N0, N1, N2);
                    X0 = N0; X1 = N1; X2 = N2; X3 = N3;
                    ++xi;
                    if (m + j + PF < nfull) {
                        const float* lp = bp + (int64_t)(4 * (m + j + PF)) * ldb;
                        q[j][0] = ldg_stream_f4(lp);
                        q[j][1] = ldg_stream_f4(lp + ldb);
                        q[j][2] = ldg_stream_f4(lp + 2 * ldb);
                        q[j][3] = ldg_stream_f4(lp + 3 * ldb);
                    }
                }
            }
        }
        for (int d = d0 + 4 * nfull; d < d1; ++d) {   // < 4 leftover diagonals
            const float4 c = ldg_stream_f4(band + (int64_t)d * ldb + row4);
            const int j = 4 * g + d;
            const double2 xa = xw[((j) & 3) * PL + ((j) >> 2)];
            const double2 xb = xw[((j + 1) & 3) * PL + ((j + 1) >> 2)];
            const double2 xc = xw[((j + 2) & 3) * PL + ((j + 2) >> 2)];
            const double2 xd = xw[((j + 3) & 3) * PL + ((j + 3) >> 2)];
            FMA4(c, xa, xb, xc, xd);
        }
    }
    // cross-segment reduction (fixed order)
    if (S > 1) {
        constexpr int G = TR / 4;
        if (s > 0) {
            double* rp = redseg + (size_t)(s - 1) * 8 * G + g;
            rp[0 * G] = acc0.x; rp[1 * G] = acc0.y; rp[2 * G] = acc1.x; rp[3 * G] = acc1.y;
            rp[4 * G] = acc2.x; rp[5 * G] = acc2.y; rp[6 * G] = acc3.x; rp[7 * G] = acc3.y;
        }
        __syncthreads();
        if (s == 0) {
#pragma unroll
            for (int ss = 1; ss < S; ++ss) {
                const double* rp = redseg + (size_t)(ss - 1) * 8 * G + g;
                acc0.x += rp[0 * G]; acc0.y += rp[1 * G]; acc1.x += rp[2 * G]; acc1.y += rp[3 * G];
                acc2.x += rp[4 * G]; acc2.y += rp[5 * G]; acc3.x += rp[6 * G]; acc3.y += rp[7 * G];
            }
        }
    }
    double dots[2] = {0.0, 0.0};
    if (s == 0 && active) {
        const int jw = 4 * g + w;   // window index of row4
        double2 accs[4] = {acc0, acc1, acc2, acc3};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int64_t i = row4 + e;
            if (i < a.M) {
                const int j = jw + e;
                epi_row<EPI>(a, i, accs[e], xw[(j & 3) * PL + (j >> 2)], dots);
            }
        }
    }
    if (EPI != EPI_PLAIN) grid_reduce<2>(dots, a.rc, red);
}

// ---------------------------------------------------------------------------------------------
// dense-panel kernel (DENSE and BLOCKDIAG layouts).  R is symmetric, so the stored row j of a
// panel is also column j: the kernel sweeps stored rows ("columns") j and accumulates
// y[i..i+3] += P[j][i..i+3] * v[j] with y in registers and v[j] a shared-memory broadcast;
// no shared-memory traffic per matrix element, fully coalesced 128-bit loads.
// Cross-CTA partial sums over j-segments go to ypart[slot][i]; k_panel_finish adds the slots in
// fixed order and applies the epilogue.
// ---------------------------------------------------------------------------------------------
#define PANEL_JC 1024

#define PFMA(C, X)                                                                            \
    do {                                                                                      \
        double v0 = (double)(C).x, v1 = (double)(C).y, v2 = (double)(C).z, v3 = (double)(C).w; \
        acc0.x = fma(v0, (X).x, acc0.x); acc0.y = fma(v0, (X).y, acc0.y);                     \
        acc1.x = fma(v1, (X).x, acc1.x); acc1.y = fma(v1, (X).y, acc1.y);                     \
        acc2.x = fma(v2, (X).x, acc2.x); acc2.y = fma(v2, (X).y, acc2.y);                     \
        acc3.x = fma(v3, (X).x, acc3.x); acc3.y = fma(v3, (X).y, acc3.y);                     \
    } while (0)

template <int RW, int S>
__global__ void __launch_bounds__(32 * RW * S, 2)
k_spmm_panel(SpmmArgs a, const float* __restrict__ panels, const PanelItem* __restrict__ items, double2* ypart,
             const double2* __restrict__ vin) {
    if (a.check_done && a.rc.st->done[0] && a.rc.st->done[1]) return;
    constexpr int TI = 128 * RW;
    constexpr int NT = 32 * RW * S;
    constexpr int G = TI / 4;
    __shared__ double2 xs[PANEL_JC];
    __shared__ double redseg[(S > 1 ? (S - 1) : 1) * 8 * G];
    const PanelItem it = items[blockIdx.x];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int rw = wid % RW, s = wid / RW;
    const int g = rw * 32 + lane;
    const bool active = 4 * g < it.navail;
    const float* base = panels + it.off + 4 * g;

    double2 acc0 = make_double2(0, 0), acc1 = acc0, acc2 = acc0, acc3 = acc0;
    for (int jb = 0; jb < it.nj; jb += PANEL_JC) {
        const int cnt = min(PANEL_JC, it.nj - jb);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt; t += NT) xs[t] = vin[(int64_t)it.j0 + jb + t];
        __syncthreads();
        const int per = (cnt + S - 1) / S;
        const int lo = s * per, hi = min(cnt, lo + per);
        if (active && lo < hi) {
            const float* p = base + (int64_t)(jb + lo) * it.ld;
            int jj = lo;
            for (; jj + 4 <= hi; jj += 4) {
                const float4 c0 = ldg_stream_f4(p);
                const float4 c1 = ldg_stream_f4(p + it.ld);
                const float4 c2 = ldg_stream_f4(p + 2 * (int64_t)it.ld);
                const float4 c3 = ldg_stream_f4(p + 3 * (int64_t)it.ld);
                const double2 x0 = xs[jj], x1 = xs[jj + 1], x2 = xs[jj + 2], x3 = xs[jj + 3];
                PFMA(c0, x0);
                PFMA(c1, x1);
                PFMA(c2, x2);
                PFMA(c3, x3);
                p += 4 * (int64_t)it.ld;
            }
            for (; jj < hi; ++jj) {
                const float4 c0 = ldg_stream_f4(p);
                const double2 x0 = xs[jj];
                PFMA(c0, x0);
                p += it.ld;
            }
        }
    }
    if (S > 1) {
        __syncthreads();
        if (s > 0) {
            double* rp = redseg + (size_t)(s - 1) * 8 * G + g;
            rp[0 * G] = acc0.x; rp[1 * G] = acc0.y; rp[2 * G] = acc1.x; rp[3 * G] = acc1.y;
            rp[4 * G] = acc2.x; rp[5 * G] = acc2.y; rp[6 * G] = acc3.x; rp[7 * G] = acc3.y;
        }
        __syncthreads();
        if (s == 0) {
#pragma unroll
            for (int ss = 1; ss < S; ++ss) {
                const double* rp = redseg + (size_t)(ss - 1) * 8 * G + g;
                acc0.x += rp[0 * G]; acc0.y += rp[1 * G]; acc1.x += rp[2 * G]; acc1.y += rp[3 * G];
                acc2.x += rp[4 * G]; acc2.y += rp[5 * G]; acc3.x += rp[6 * G]; acc3.y += rp[7 * G];
            }
        }
    }
    if (s == 0) {
        double2* yp = ypart + (int64_t)it.slot * a.M + it.i0 + 4 * g;
        const int left = it.ni - 4 * g;
        if (left > 0) yp[0] = acc0;
        if (left > 1) yp[1] = acc1;
        if (left > 2) yp[2] = acc2;
        if (left > 3) yp[3] = acc3;
    }
}

// Rows partition of a dense R (ld.rowpart): the input vector pair of ALL ranks, read from the peers' symmetric arenas
// into this rank's `vfull` (M entries, 16 B each: 1.6 MB at M = 100k against the 5 GB panel the product then streams).
// fused: the CG direction update is formed on the way, p_new = r + beta p_old (scipy: p *= beta; p += z) from every
// rank's r and p_old - the same values every owner computes - and the own slice of p_new is stored; p is double-
// buffered, so a peer still gathering p_old is never overtaken.  Ordering: r, p_old and x of every rank are complete
// once the previous cross-rank reduction has resolved (k_cg_update / k_lmmse_setup / k_pack_x0), as for the halos of
// the banded partition.
struct GatherArgs {
    const double2* v[SGV_MAX_RANKS];
    const double2* r[SGV_MAX_RANKS];
    int64_t        lo[SGV_MAX_RANKS + 1];
    int            world, rank;
};
__global__ void __launch_bounds__(256)
k_gather_rows(GatherArgs g, double2* __restrict__ vfull, double2* __restrict__ p_new, int fused, int check_done,
              const CgState* __restrict__ st) {
    const int d0 = st->done[0], d1 = st->done[1];
    if (check_done && d0 && d1) return;
    const bool first = st->step == 0;
    double b0 = 0.0, b1 = 0.0;
    if (fused && !first) {
        b0 = st->rho[0] / st->rho_prev[0];
        b1 = st->rho[1] / st->rho_prev[1];
    }
    const int64_t M = g.lo[g.world];
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < M; t += (int64_t)gridDim.x * blockDim.x) {
        int q = 0;
        while (q + 1 < g.world && t >= g.lo[q + 1]) ++q;
        const int64_t i = t - g.lo[q];
        double2 p;
        if (fused) {
            const double2 r = ld_vec2(g.r[q] + i);
            p = first ? make_double2(0.0, 0.0) : ld_vec2(g.v[q] + i);
            if (!d0) p.x = first ? r.x : (p.x * b0 + r.x);
            if (!d1) p.y = first ? r.y : (p.y * b1 + r.y);
            if (q == g.rank) p_new[i] = p;
        } else {
            p = ld_vec2(g.v[q] + i);
        }
        vfull[t] = p;
    }
}

template <int EPI>
__global__ void __launch_bounds__(256) k_panel_finish(SpmmArgs a, const double2* __restrict__ ypart, int nslots) {
    SGV_LOAD_DEV_SCALARS(a);
    if (a.check_done && a.rc.st->done[0] && a.rc.st->done[1]) return;
    __shared__ double red[2 * 32];
    double dots[2] = {0.0, 0.0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.M; i += (int64_t)gridDim.x * blockDim.x) {
        double2 acc = ypart[i];
        for (int sl = 1; sl < nslots; ++sl) {
            const double2 t = ypart[(int64_t)sl * a.M + i];
            acc.x += t.x;
            acc.y += t.y;
        }
        epi_row<EPI>(a, i, acc, a.v[i], dots);
    }
    if (EPI != EPI_PLAIN) grid_reduce<2>(dots, a.rc, red);
}

// ---------------------------------------------------------------------------------------------
// CSR kernel (general sparsity): one warp per row, coalesced value / index loads, 16-byte gathers
// of the vector pair through the read-only path (the vectors stay L2 resident).
// ---------------------------------------------------------------------------------------------
template <int EPI>
__global__ void __launch_bounds__(256)
k_spmm_csr(SpmmArgs a, const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
           const float* __restrict__ vals, const double2* __restrict__ vin) {
    SGV_LOAD_DEV_SCALARS(a);
    if (a.check_done && a.rc.st->done[0] && a.rc.st->done[1]) return;
    __shared__ double red[2 * 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double dots[2] = {0.0, 0.0};
    // grid-stride over rows: a bounded number of blocks (one ticket atomic and one partial row each)
    for (int64_t row = (int64_t)blockIdx.x * 8 + wid; row < a.M; row += (int64_t)gridDim.x * 8) {
        const int64_t beg = indptr[row], end = indptr[row + 1];
        double ax = 0.0, ay = 0.0, bx = 0.0, by = 0.0;
        int64_t k = beg + lane;
        for (; k + 32 < end; k += 64) {
            const float v0 = ldg_stream_f1(vals + k), v1 = ldg_stream_f1(vals + k + 32);
            const int c0 = ldg_stream_i1(indices + k), c1 = ldg_stream_i1(indices + k + 32);
            const double2 x0 = __ldg(&vin[c0]);
            const double2 x1 = __ldg(&vin[c1]);
            ax = fma((double)v0, x0.x, ax); ay = fma((double)v0, x0.y, ay);
            bx = fma((double)v1, x1.x, bx); by = fma((double)v1, x1.y, by);
        }
        if (k < end) {
            const float v0 = ldg_stream_f1(vals + k);
            const int c0 = ldg_stream_i1(indices + k);
            const double2 x0 = __ldg(&vin[c0]);
            ax = fma((double)v0, x0.x, ax); ay = fma((double)v0, x0.y, ay);
        }
        ax = warp_sum(ax + bx);
        ay = warp_sum(ay + by);
        if (lane == 0) epi_row<EPI>(a, row, make_double2(ax, ay), a.v[row], dots);
    }
    if (EPI != EPI_PLAIN) grid_reduce<2>(dots, a.rc, red);
}

// ---------------------------------------------------------------------------------------------
// host-side dispatch
// ---------------------------------------------------------------------------------------------
template <int RW, int S, int EPI, int PF, int MINB>
static int launch_dia(sgv_ctx* c, const LdMatrix& ld, SpmmArgs& a) {
    constexpr int TR = 128 * RW;
    const size_t smem = sgv_dia_smem_bytes(ld.w, RW, S);
    const unsigned grid = (unsigned)((ld.ldb + TR - 1) / TR);
    SGV_TRY(sgv_ensure_partials(c, grid));
    a.rc.partials = c->partials;
    k_spmm_dia<RW, S, EPI, PF, MINB><<<grid, 32 * RW * S, smem, c->stream>>>(a, ld.band, (int)ld.w, ld.ldb);
    c->launches++;
    return 0;
}

// default tile shapes
#define DIA_BIG_RW 2
#define DIA_BIG_S 4
#define DIA_PF 2
#define DIA_MINB 3

template <int EPI>
static int launch_epi(sgv_ctx* c, Cohort& co, SpmmArgs& a) {
    const LdMatrix& ld = co.ld;
    if (ld.layout == SGV_LAYOUT_DIA) {
        // wide tiles when there are enough rows to fill the machine, narrow ones otherwise
        const bool big = ld.ldb >= (int64_t)c->sm_count * 2 * 256 * 2;
        if (big && sgv_dia_smem_bytes(ld.w, DIA_BIG_RW, DIA_BIG_S) <= 72 * 1024)
            return launch_dia<DIA_BIG_RW, DIA_BIG_S, EPI, DIA_PF, DIA_MINB>(c, ld, a);
        return launch_dia<1, 8, EPI, DIA_PF, DIA_MINB>(c, ld, a);
    }
    if (ld.layout == SGV_LAYOUT_DSYM) return sgv_launch_dsym(c, ld, EPI, a);
    // Default for dense / block-diagonal LD: the upper-triangle kernel of spmm_psym.cu (half the bytes read).
    // Measured on B200 (2-RHS pass): dense M=50k 1.10 ms vs 1.53 ms for the full-panel kernel below (which runs at
    // the HBM roofline of the FULL matrix), M=10k 0.070 vs 0.086 ms, block-diagonal M=300k 0.45 vs 0.52 ms.
    // SGV_PANEL_FULL=1 selects the full-panel kernel.
    if ((ld.layout == SGV_LAYOUT_DENSE || ld.layout == SGV_LAYOUT_BLOCKDIAG) && ld.sym_items != nullptr && ld.panel_sym) {
        const char* e = getenv("SGV_PANEL_FULL");
        if (e == nullptr || e[0] != '1') return sgv_launch_psym(c, ld, EPI, a);
    }
    if (ld.layout == SGV_LAYOUT_DENSE || ld.layout == SGV_LAYOUT_BLOCKDIAG) {
        k_spmm_panel<4, 2><<<ld.n_items, 256, 0, c->stream>>>(a, ld.panels, ld.items, c->ypart, ld.rowpart ? c->vfull : a.v);
        c->launches++;
        const unsigned grid = (unsigned)std::min<int64_t>((c->Ml + 255) / 256, (int64_t)c->sm_count * 6);
        SGV_TRY(sgv_ensure_partials(c, grid));
        a.rc.partials = c->partials;
        k_panel_finish<EPI><<<grid, 256, 0, c->stream>>>(a, c->ypart, ld.s_cross);
        c->launches++;
        return 0;
    }
    if (ld.layout == SGV_LAYOUT_CSR) {
        const unsigned grid = (unsigned)std::min<int64_t>((c->Ml + 7) / 8, (int64_t)c->sm_count * 32);
        SGV_TRY(sgv_ensure_partials(c, grid));
        a.rc.partials = c->partials;
        k_spmm_csr<EPI><<<grid, 256, 0, c->stream>>>(a, ld.indptr, ld.indices, ld.vals, ld.rowpart ? c->vfull : a.v);
        c->launches++;
        return 0;
    }
    sgv_set_error("cohort has no LD matrix uploaded");
    return -1;
}

// Load every SpMM kernel on the current device and raise its dynamic shared-memory limit once, at
// handle creation.  Lazy module loading and cudaFuncSetAttribute synchronise the device; done inside
// the solver loop they would deadlock against another rank's resolve kernel spinning on the same GPU.
#define DIA_SMEM_LIMIT (200 * 1024)
template <int RW, int S, int EPI>
static int preload_dia() {
    SGV_CUDA(cudaFuncSetAttribute(k_spmm_dia<RW, S, EPI, DIA_PF, DIA_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  DIA_SMEM_LIMIT));
    return 0;
}
template <int EPI>
static int preload_epi() {
    SGV_TRY((preload_dia<DIA_BIG_RW, DIA_BIG_S, EPI>()));
    SGV_TRY((preload_dia<1, 8, EPI>()));
    cudaFuncAttributes fa;
    SGV_CUDA(cudaFuncGetAttributes(&fa, k_panel_finish<EPI>));
    SGV_CUDA(cudaFuncGetAttributes(&fa, k_spmm_csr<EPI>));
    return 0;
}
int sgv_preload_spmm() {
    SGV_TRY(preload_epi<EPI_Q>());
    SGV_TRY(preload_epi<EPI_RESID>());
    SGV_TRY(preload_epi<EPI_STATS>());
    SGV_TRY(preload_epi<EPI_PLAIN>());
    cudaFuncAttributes fa;
    SGV_CUDA(cudaFuncGetAttributes(&fa, k_spmm_panel<4, 2>));
    SGV_CUDA(cudaFuncGetAttributes(&fa, k_gather_rows));
    return 0;
}

static double2* arena_vec(const sgv_ctx* c, int q, int vec) {
    const PeerView& pv = c->peer[q];
    const size_t off = vec == VEC_XX ? arena_off_xx(pv.Ml) : arena_off_pp(pv.Ml, vec - VEC_PP0);
    return reinterpret_cast<double2*>(pv.base + off);
}

// vec: which symmetric vector is the SpMM input (VEC_XX / VEC_PP0 / VEC_PP1).  With fused_p it names
// p_old; the new direction r + beta*p_old is written to the other pp buffer.
int sgv_launch_spmm(sgv_ctx* c, Cohort& co, int epi, int vec, double2* out, double gamw, double gam2, int check_done,
                    int fused_p) {
    SpmmArgs a;
    memset(&a, 0, sizeof(a));
    a.vs_cohort = c->vs_active;   // >= 0 inside the fused VAMP iteration: gamw / gam2 come from the device
    a.v = vec == VEC_XX ? c->xx : c->pp[vec - VEC_PP0];
    a.fused_p = fused_p;
    const bool rowpart = co.ld.rowpart;   // DENSE column panel or CSR rows with global column indices
    SGV_CHECK(!c->rowpart || rowpart, "a handle configured for the rows partition holds dense column panels or CSR rows only");
    if (fused_p) {
        SGV_CHECK((co.ld.layout == SGV_LAYOUT_DIA || rowpart) && vec != VEC_XX,
                  "fused direction update needs the DIA layout or the rows partition");
        a.r = c->rr;
        a.p_new = c->pp[1 - (vec - VEC_PP0)];
    }
    if (c->world > 1 && c->halo && (co.ld.layout == SGV_LAYOUT_DIA || co.ld.layout == SGV_LAYOUT_DSYM)) {
        if (c->rank > 0) {
            const PeerView& pv = c->peer[c->rank - 1];
            SGV_CHECK(pv.base != nullptr && pv.Ml >= co.ld.w, "left neighbour not attached or shorter than the half-bandwidth");
            a.v_left = arena_vec(c, c->rank - 1, vec);
            a.r_left = reinterpret_cast<double2*>(pv.base + arena_off_rr(pv.Ml, 1));
            a.n_left = pv.Ml;
        }
        if (c->rank + 1 < c->world) {
            const PeerView& pv = c->peer[c->rank + 1];
            SGV_CHECK(pv.base != nullptr && pv.Ml >= co.ld.w, "right neighbour not attached or shorter than the half-bandwidth");
            a.v_right = arena_vec(c, c->rank + 1, vec);
            a.r_right = reinterpret_cast<double2*>(pv.base + arena_off_rr(pv.Ml, 1));
        }
    }
    a.out = out;
    a.bb = c->bb;
    a.gamw = gamw;
    a.gam2 = gam2;
    a.M = c->Ml;
    a.check_done = check_done;
    const int kind = epi == EPI_Q ? AP_PQ : epi == EPI_RESID ? AP_RESID : AP_STATS;
    if (epi != EPI_PLAIN) {
        a.rc = sgv_red_begin(c, kind, 2, 0);
        a.rc.skip_if_done = check_done ? SKIP_CG_DONE : SKIP_NEVER;
    }
    a.rc.st = c->cg;
    int rc;
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
    if (rowpart) {
        // all-gather of the input vector pair (with the direction update when fused); the epilogue then works on
        // the own slice: p_new when fused
        GatherArgs g;
        memset(&g, 0, sizeof(g));
        g.world = c->world;
        g.rank = c->rank;
        int64_t lo = 0;
        for (int q = 0; q < c->world; ++q) {
            const PeerView& pv = c->peer[q];
            SGV_CHECK(pv.base != nullptr, "rank %d not attached (rows partition reads every rank's vectors)", q);
            g.v[q] = arena_vec(c, q, vec);
            g.r[q] = reinterpret_cast<const double2*>(pv.base + arena_off_rr(pv.Ml, 1));
            g.lo[q] = lo;
            if (q == c->rank) SGV_CHECK(lo == c->row_lo, "rows partition must be contiguous in rank order");
            lo += pv.Ml;
        }
        g.lo[c->world] = lo;
        SGV_CHECK(lo == c->M, "the ranks' row ranges do not add up to M");
        const unsigned ggrid = (unsigned)std::min<int64_t>((c->M + 255) / 256, (int64_t)c->sm_count * 8);
        k_gather_rows<<<ggrid, 256, 0, c->stream>>>(g, c->vfull, a.p_new, fused_p, check_done, c->cg);
        c->launches++;
        if (fused_p) a.v = a.p_new;
        a.fused_p = 0;
    }
    switch (epi) {
        case EPI_Q: rc = launch_epi<EPI_Q>(c, co, a); break;
        case EPI_RESID: rc = launch_epi<EPI_RESID>(c, co, a); break;
        case EPI_STATS: rc = launch_epi<EPI_STATS>(c, co, a); break;
        default: rc = launch_epi<EPI_PLAIN>(c, co, a); break;
    }
    if (rc) return rc;
    if (c->prof) {
        SGV_CUDA(cudaEventRecord(c->prof_ev[c->prof_n + 1], c->stream));
        c->prof_n += 2;
    }
    SGV_CUDA(cudaGetLastError());
    if (epi != EPI_PLAIN) SGV_TRY(sgv_red_end(c, a.rc));
    return 0;
}
