// Dense-panel SpMM that reads only the UPPER triangle of every (symmetric) panel: out = gamw*(R v) + gam2*v
// for a pair of fp64 vectors, R dense or block-diagonal with dense blocks, stored row-major fp32 (the same
// buffers as the full-panel kernel of spmm.cu - nothing is stored differently, half of it is simply not read).
//
// A work item is a strip of TI = 512 columns [i0, i0+ni) of one panel times a segment of stored rows
// [j0, j0+nj) with j0 + nj <= i0 + ni (on or above the diagonal).  A thread owns 4 consecutive columns and
// sweeps the rows; every float4 it loads is used twice from registers:
//   forward     y[i..i+3] += P[j][i..i+3] * v[j]           y in registers, v[j] a shared-memory broadcast
//   transposed  y[j]      += P[j][i..i+3] . v[i..i+3]      v[i..i+3] fixed in registers; the 4 rows x 2 RHS
//               partial dots of a group of 4 rows are summed over the warp with a transposing butterfly
//               (9 double shuffles instead of 40) and land in a per-row-warp shared array.
// Rows inside the strip's own diagonal tile use masks (forward i >= j, transposed i > j).  The forward
// partial of segment `slot` goes to ypart[slot][i], the transposed partial of strip `strip` to
// ypartT[strip][j]; k_psym_finish adds, in fixed order, the slots of a row's strip and the strips at or right of
// it, and applies the fused epilogue + reduction.  No atomics; bit-reproducible.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include "sgv_device.cuh"

#define PS_RW 4
#define PS_S 2
#define PS_TI (128 * PS_RW)
#define PS_JC 512

struct SymItem {
    int64_t off;      // element offset of P[j0][i0]
    int     ld;
    int     i0, ni;   // global column range of the strip
    int     j0, nj;   // global stored-row range of the segment
    int     navail;   // readable floats from column i0 to the end of the padded row
    int     slot;     // forward slot (segment index within the strip)
    int     strip;    // strip index within its block
    int     diag;     // rows reach into [i0, i0+ni): masks needed
};

#define PS_FWD(C, X)                                                                          \
    do {                                                                                      \
        acc0.x = fma((double)(C).x, (X).x, acc0.x); acc0.y = fma((double)(C).x, (X).y, acc0.y); \
        acc1.x = fma((double)(C).y, (X).x, acc1.x); acc1.y = fma((double)(C).y, (X).y, acc1.y); \
        acc2.x = fma((double)(C).z, (X).x, acc2.x); acc2.y = fma((double)(C).z, (X).y, acc2.y); \
        acc3.x = fma((double)(C).w, (X).x, acc3.x); acc3.y = fma((double)(C).w, (X).y, acc3.y); \
    } while (0)

__device__ __forceinline__ double2 ps_dot(const float4 c, const double2 O0, const double2 O1, const double2 O2, const double2 O3) {
    double2 d;
    d.x = (double)c.x * O0.x;
    d.y = (double)c.x * O0.y;
    d.x = fma((double)c.y, O1.x, d.x); d.y = fma((double)c.y, O1.y, d.y);
    d.x = fma((double)c.z, O2.x, d.x); d.y = fma((double)c.z, O2.y, d.y);
    d.x = fma((double)c.w, O3.x, d.x); d.y = fma((double)c.w, O3.y, d.y);
    return d;
}

__device__ __forceinline__ float4 ps_mask(float4 c, int i, int j, bool strict) {
    // keep element e iff column i+e is right of (strict) / on-or-right of row j
    const int t = strict ? j + 1 : j;
    if (i < t) c.x = 0.f;
    if (i + 1 < t) c.y = 0.f;
    if (i + 2 < t) c.z = 0.f;
    if (i + 3 < t) c.w = 0.f;
    return c;
}

__global__ void __launch_bounds__(32 * PS_RW * PS_S, 2)
k_spmm_psym(SpmmArgs a, const float* __restrict__ panels, const SymItem* __restrict__ items, double2* __restrict__ ypart,
            double2* __restrict__ ypartT) {
    if (a.check_done && a.rc.st->done[0] && a.rc.st->done[1]) return;
    constexpr int NT = 32 * PS_RW * PS_S;
    constexpr int G = PS_TI / 4;
    extern __shared__ double2 ps_smem[];
    double2* xs = ps_smem;                                   // v[j] of the current chunk
    double2* yTw = xs + PS_JC;                               // [PS_RW][PS_JC] transposed partials per row-warp
    double* redseg = reinterpret_cast<double*>(yTw + PS_RW * PS_JC);   // (S-1) * 8 * G
    const SymItem it = items[blockIdx.x];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int rw = wid % PS_RW, s = wid / PS_RW;
    const int g = rw * 32 + lane;
    const bool active = 4 * g < it.navail;
    const float* base = panels + it.off + 4 * g;
    const int icol = it.i0 + 4 * g;                          // first of this thread's 4 columns
    const int iend = it.i0 + it.ni;
    const double2 zero2 = make_double2(0.0, 0.0);
    const double2 O0 = icol < iend ? a.v[icol] : zero2, O1 = icol + 1 < iend ? a.v[icol + 1] : zero2,
                  O2 = icol + 2 < iend ? a.v[icol + 2] : zero2, O3 = icol + 3 < iend ? a.v[icol + 3] : zero2;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    double2* yT = yTw + rw * PS_JC;

    double2 acc0 = zero2, acc1 = zero2, acc2 = zero2, acc3 = zero2;
    for (int jb = 0; jb < it.nj; jb += PS_JC) {
        const int cnt = min(PS_JC, it.nj - jb);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt; t += NT) xs[t] = a.v[(int64_t)it.j0 + jb + t];
        __syncthreads();
        const int per = (((cnt + PS_S - 1) / PS_S) + 3) & ~3;
        const int lo = min(cnt, s * per), hi = min(cnt, lo + per);
        const float* p = base + (int64_t)(jb + lo) * it.ld;
        // software pipeline: the 4 rows of the next group are in flight while the current group is multiplied
        float4 n0 = zero4, n1 = zero4, n2 = zero4, n3 = zero4;
        if (active && lo < hi) {
            const int nr = min(4, hi - lo);
            n0 = ldg_stream_f4(p);
            if (nr > 1) n1 = ldg_stream_f4(p + it.ld);
            if (nr > 2) n2 = ldg_stream_f4(p + 2 * (int64_t)it.ld);
            if (nr > 3) n3 = ldg_stream_f4(p + 3 * (int64_t)it.ld);
        }
        for (int jj = lo; jj < hi; jj += 4) {
            float4 c0 = n0, c1 = n1, c2 = n2, c3 = n3;
            p += 4 * (int64_t)it.ld;
            n0 = n1 = n2 = n3 = zero4;
            if (active && jj + 4 < hi) {
                const int nr = min(4, hi - (jj + 4));        // warp-uniform
                n0 = ldg_stream_f4(p);
                if (nr > 1) n1 = ldg_stream_f4(p + it.ld);
                if (nr > 2) n2 = ldg_stream_f4(p + 2 * (int64_t)it.ld);
                if (nr > 3) n3 = ldg_stream_f4(p + 3 * (int64_t)it.ld);
            }
            const int jg = it.j0 + jb + jj;                  // global row of c0
            const double2 x0 = xs[jj], x1 = xs[min(jj + 1, cnt - 1)], x2 = xs[min(jj + 2, cnt - 1)], x3 = xs[min(jj + 3, cnt - 1)];
            double2 d0, d1, d2, d3;
            if (it.diag && jg + 3 >= it.i0) {                // inside the strip's diagonal tile (warp-uniform): masked copies
                const float4 t0 = ps_mask(c0, icol, jg, true), t1 = ps_mask(c1, icol, jg + 1, true),
                             t2 = ps_mask(c2, icol, jg + 2, true), t3 = ps_mask(c3, icol, jg + 3, true);
                c0 = ps_mask(c0, icol, jg, false);
                c1 = ps_mask(c1, icol, jg + 1, false);
                c2 = ps_mask(c2, icol, jg + 2, false);
                c3 = ps_mask(c3, icol, jg + 3, false);
                PS_FWD(c0, x0);
                PS_FWD(c1, x1);
                PS_FWD(c2, x2);
                PS_FWD(c3, x3);
                d0 = ps_dot(t0, O0, O1, O2, O3);
                d1 = ps_dot(t1, O0, O1, O2, O3);
                d2 = ps_dot(t2, O0, O1, O2, O3);
                d3 = ps_dot(t3, O0, O1, O2, O3);
            } else {
                // the common case: both uses read the SAME registers, so every value is converted to fp64 once
                // (with separate masked copies the compiler emitted two F2F per value: XU pipe 53 % -> see DESIGN 4)
                PS_FWD(c0, x0);
                d0 = ps_dot(c0, O0, O1, O2, O3);
                PS_FWD(c1, x1);
                d1 = ps_dot(c1, O0, O1, O2, O3);
                PS_FWD(c2, x2);
                d2 = ps_dot(c2, O0, O1, O2, O3);
                PS_FWD(c3, x3);
                d3 = ps_dot(c3, O0, O1, O2, O3);
            }
            // transposing butterfly: 8 values over 32 lanes -> lane L (L % 4 == 0) holds the warp sum of value L / 4
            double v0 = d0.x, v1 = d0.y, v2 = d1.x, v3 = d1.y, v4 = d2.x, v5 = d2.y, v6 = d3.x, v7 = d3.y;
            {
                const bool up = (lane & 16) != 0;
                const double s0 = up ? v0 : v4, s1 = up ? v1 : v5, s2 = up ? v2 : v6, s3 = up ? v3 : v7;
                const double k0 = up ? v4 : v0, k1 = up ? v5 : v1, k2 = up ? v6 : v2, k3 = up ? v7 : v3;
                v0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 16);
                v1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 16);
                v2 = k2 + __shfl_xor_sync(0xffffffffu, s2, 16);
                v3 = k3 + __shfl_xor_sync(0xffffffffu, s3, 16);
            }
            {
                const bool up = (lane & 8) != 0;
                const double s0 = up ? v0 : v2, s1 = up ? v1 : v3;
                const double k0 = up ? v2 : v0, k1 = up ? v3 : v1;
                v0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 8);
                v1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 8);
            }
            {
                const bool up = (lane & 4) != 0;
                const double s0 = up ? v0 : v1;
                const double k0 = up ? v1 : v0;
                v0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 4);
            }
            v0 += __shfl_xor_sync(0xffffffffu, v0, 2);
            v0 += __shfl_xor_sync(0xffffffffu, v0, 1);
            if ((lane & 3) == 0) {
                const int vid = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);   // row = vid / 2, rhs = vid % 2
                const int row = jj + (vid >> 1);
                if (row < hi) reinterpret_cast<double*>(yT + row)[vid & 1] = v0;
            }
        }
        __syncthreads();
        // transposed partial of this strip for the chunk's rows: the PS_RW row-warps in fixed order
        for (int t = threadIdx.x; t < cnt; t += NT) {
            double2 sum = yTw[t];
#pragma unroll
            for (int r2 = 1; r2 < PS_RW; ++r2) {
                const double2 q = yTw[r2 * PS_JC + t];
                sum.x += q.x;
                sum.y += q.y;
            }
            ypartT[(int64_t)it.strip * a.M + it.j0 + jb + t] = sum;
        }
    }
    // forward sums: the S segments of the CTA in fixed order
    if (PS_S > 1) {
        __syncthreads();
        if (s > 0) {
            double* rp = redseg + (size_t)(s - 1) * 8 * G + g;
            rp[0 * G] = acc0.x; rp[1 * G] = acc0.y; rp[2 * G] = acc1.x; rp[3 * G] = acc1.y;
            rp[4 * G] = acc2.x; rp[5 * G] = acc2.y; rp[6 * G] = acc3.x; rp[7 * G] = acc3.y;
        }
        __syncthreads();
        if (s == 0) {
#pragma unroll
            for (int ss = 1; ss < PS_S; ++ss) {
                const double* rp = redseg + (size_t)(ss - 1) * 8 * G + g;
                acc0.x += rp[0 * G]; acc0.y += rp[1 * G]; acc1.x += rp[2 * G]; acc1.y += rp[3 * G];
                acc2.x += rp[4 * G]; acc2.y += rp[5 * G]; acc3.x += rp[6 * G]; acc3.y += rp[7 * G];
            }
        }
    }
    if (s == 0) {
        double2* yp = ypart + (int64_t)it.slot * a.M + icol;
        const int left = iend - icol;
        if (left > 0) yp[0] = acc0;
        if (left > 1) yp[1] = acc1;
        if (left > 2) yp[2] = acc2;
        if (left > 3) yp[3] = acc3;
    }
}

// rowmeta[3*i + {0,1,2}] = strip of row i within its block, number of strips of the block, forward slots of the strip
template <int EPI>
__global__ void __launch_bounds__(256)
k_psym_finish(SpmmArgs a, const double2* __restrict__ ypart, const double2* __restrict__ ypartT, const int* __restrict__ rowmeta) {
    SGV_LOAD_DEV_SCALARS(a);
    if (a.check_done && a.rc.st->done[0] && a.rc.st->done[1]) return;
    __shared__ double red[2 * 32];
    double dots[2] = {0.0, 0.0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.M; i += (int64_t)gridDim.x * blockDim.x) {
        const int strip = rowmeta[3 * i], nstrips = rowmeta[3 * i + 1], nseg = rowmeta[3 * i + 2];
        double2 y = make_double2(0.0, 0.0);
        for (int sl = 0; sl < nseg; ++sl) {
            const double2 t = ypart[(int64_t)sl * a.M + i];
            y.x += t.x;
            y.y += t.y;
        }
        for (int st = strip; st < nstrips; ++st) {
            const double2 t = ypartT[(int64_t)st * a.M + i];
            y.x += t.x;
            y.y += t.y;
        }
        epi_row<EPI>(a, i, y, a.v[i], dots);
    }
    if (EPI != EPI_PLAIN) grid_reduce<2>(dots, a.rc, red);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static size_t psym_smem() {
    return (size_t)(PS_JC + PS_RW * PS_JC) * sizeof(double2) + (size_t)(PS_S - 1) * 8 * (PS_TI / 4) * sizeof(double);
}

int sgv_preload_psym() {
    SGV_CUDA(cudaFuncSetAttribute(k_spmm_psym, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psym_smem()));
    cudaFuncAttributes fa;
    SGV_CUDA(cudaFuncGetAttributes(&fa, k_psym_finish<EPI_Q>));
    SGV_CUDA(cudaFuncGetAttributes(&fa, k_psym_finish<EPI_RESID>));
    SGV_CUDA(cudaFuncGetAttributes(&fa, k_psym_finish<EPI_STATS>));
    SGV_CUDA(cudaFuncGetAttributes(&fa, k_psym_finish<EPI_PLAIN>));
    return 0;
}

// starts / offs / lds: the dense blocks of the layout (see sgv_build_panel_items)
int sgv_build_psym_items(sgv_ctx* c, LdMatrix& ld, const std::vector<int64_t>& starts, const std::vector<int64_t>& offs,
                         const std::vector<int>& lds) {
    const int nb = (int)starts.size() - 1;
    const int64_t M = c->Ml;
    double total_rows = 0;
    for (int b = 0; b < nb; ++b) {
        const int64_t m = starts[b + 1] - starts[b];
        for (int64_t i0 = 0; i0 < m; i0 += PS_TI) total_rows += (double)std::min<int64_t>(m, i0 + PS_TI);
    }
    const int64_t target = (int64_t)c->sm_count * 6;
    int64_t L = (int64_t)(total_rows / (double)target) + 1;
    L = std::max<int64_t>(256, round_up(L, 4));
    std::vector<SymItem> items;
    std::vector<int> rowmeta(3 * (size_t)M);
    int max_slots = 1, max_strips = 1;
    for (int b = 0; b < nb; ++b) {
        const int64_t s0 = starts[b], m = starts[b + 1] - s0;
        const int nstrips = (int)((m + PS_TI - 1) / PS_TI);
        max_strips = std::max(max_strips, nstrips);
        for (int k = 0; k < nstrips; ++k) {
            const int64_t i0 = (int64_t)k * PS_TI, ni = std::min<int64_t>(PS_TI, m - i0), nrows = i0 + ni;
            const int nseg = (int)((nrows + L - 1) / L);
            max_slots = std::max(max_slots, nseg);
            for (int sg = 0; sg < nseg; ++sg) {
                SymItem it;
                const int64_t j0 = (int64_t)sg * L, nj = std::min<int64_t>(L, nrows - j0);
                it.ld = lds[b];
                it.off = offs[b] + j0 * lds[b] + i0;
                it.i0 = (int)(s0 + i0);
                it.ni = (int)ni;
                it.j0 = (int)(s0 + j0);
                it.nj = (int)nj;
                it.navail = (int)std::min<int64_t>(PS_TI, lds[b] - i0);
                it.slot = sg;
                it.strip = k;
                it.diag = (j0 + nj > i0) ? 1 : 0;
                items.push_back(it);
            }
            for (int64_t i = i0; i < i0 + ni; ++i) {
                rowmeta[3 * (size_t)(s0 + i)] = k;
                rowmeta[3 * (size_t)(s0 + i) + 1] = nstrips;
                rowmeta[3 * (size_t)(s0 + i) + 2] = nseg;
            }
        }
    }
    std::stable_sort(items.begin(), items.end(), [](const SymItem& x, const SymItem& y) { return x.nj > y.nj; });
    if (ld.sym_items) cudaFree(ld.sym_items);
    if (ld.rowmeta) cudaFree(ld.rowmeta);
    ld.sym_items = nullptr;
    ld.rowmeta = nullptr;
    SGV_CUDA(cudaMalloc(&ld.sym_items, items.size() * sizeof(SymItem)));
    SGV_CUDA(cudaMemcpy(ld.sym_items, items.data(), items.size() * sizeof(SymItem), cudaMemcpyHostToDevice));
    SGV_CUDA(cudaMalloc(&ld.rowmeta, rowmeta.size() * sizeof(int)));
    SGV_CUDA(cudaMemcpy(ld.rowmeta, rowmeta.data(), rowmeta.size() * sizeof(int), cudaMemcpyHostToDevice));
    ld.n_sym_items = (int)items.size();
    const int64_t need = (int64_t)max_slots * M, needT = (int64_t)max_strips * M;
    if (c->ypart_cap < need) {
        SGV_CUDA(cudaStreamSynchronize(c->stream));
        if (c->ypart) cudaFree(c->ypart);
        c->ypart = nullptr;
        c->ypart_cap = 0;
        SGV_CUDA(cudaMalloc(&c->ypart, need * sizeof(double2)));
        c->ypart_cap = need;
    }
    if (c->ypartT_cap < needT) {
        SGV_CUDA(cudaStreamSynchronize(c->stream));
        if (c->ypartT) cudaFree(c->ypartT);
        c->ypartT = nullptr;
        c->ypartT_cap = 0;
        SGV_CUDA(cudaMalloc(&c->ypartT, needT * sizeof(double2)));
        c->ypartT_cap = needT;
    }
    return 0;
}

template <int EPI>
static int launch_fin(sgv_ctx* c, const LdMatrix& ld, SpmmArgs& a) {
    const unsigned grid = (unsigned)std::min<int64_t>((c->Ml + 255) / 256, (int64_t)c->sm_count * 6);
    SGV_TRY(sgv_ensure_partials(c, grid));
    a.rc.partials = c->partials;
    k_psym_finish<EPI><<<grid, 256, 0, c->stream>>>(a, c->ypart, c->ypartT, ld.rowmeta);
    c->launches++;
    return 0;
}

int sgv_launch_psym(sgv_ctx* c, const LdMatrix& ld, int epi, SpmmArgs& a) {
    SGV_CHECK(ld.sym_items != nullptr && c->ypartT != nullptr, "symmetric panel items not built");
    k_spmm_psym<<<ld.n_sym_items, 32 * PS_RW * PS_S, psym_smem(), c->stream>>>(a, ld.panels, ld.sym_items, c->ypart, c->ypartT);
    c->launches++;
    switch (epi) {
        case EPI_Q: return launch_fin<EPI_Q>(c, ld, a);
        case EPI_RESID: return launch_fin<EPI_RESID>(c, ld, a);
        case EPI_STATS: return launch_fin<EPI_STATS>(c, ld, a);
        default: return launch_fin<EPI_PLAIN>(c, ld, a);
    }
}
