// LD matrix ingestion: host CSR / dense -> fp32 HBM layouts (DIA band, dense panels, CSR), with
// Rused = (1-s) R + s I (reference src/main.py:265) applied in fp64 before rounding to fp32.
// Layout detection (bandwidth, diagonal blocks, fill) runs on per-row column extents computed on
// the device, so the host never walks the nnz arrays.
#include <algorithm>
#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include "sgv_device.cuh"

void sgv_ld_free(LdMatrix& ld) {
    if (ld.owned) {
        if (ld.band) cudaFree(const_cast<float*>(ld.band));
        if (ld.panels) cudaFree(const_cast<float*>(ld.panels));
    }
    if (ld.items) cudaFree(ld.items);
    if (ld.sym_items) cudaFree(ld.sym_items);
    if (ld.rowmeta) cudaFree(ld.rowmeta);
    if (ld.indptr) cudaFree(ld.indptr);
    if (ld.indices) cudaFree(ld.indices);
    if (ld.vals) cudaFree(ld.vals);
    ld = LdMatrix();
}

int sgv_ensure_stage(sgv_ctx* c, int64_t bytes) {
    if (c->stage_bytes >= bytes) return 0;
    if (c->stage) cudaFree(c->stage);
    c->stage = nullptr;
    c->stage_bytes = 0;
    SGV_CUDA(cudaMalloc(&c->stage, bytes));
    c->stage_bytes = bytes;
    return 0;
}

int sgv_ensure_partials(sgv_ctx* c, int64_t nblocks) {
    const int64_t need = nblocks * SGV_MAX_PARTIAL_VALUES;
    if (c->partials_cap >= need) return 0;
    SGV_CUDA(cudaStreamSynchronize(c->stream));
    if (c->partials) cudaFree(c->partials);
    c->partials = nullptr;
    c->partials_cap = 0;
    SGV_CUDA(cudaMalloc(&c->partials, need * sizeof(double)));
    c->partials_cap = need;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
__global__ void k_row_extent(int64_t M, const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                             int* __restrict__ lo, int* __restrict__ hi, int* __restrict__ has_diag, int col_base) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= M) return;
    int mn = INT_MAX, mx = -1, dg = 0;
    for (int64_t k = indptr[row] + lane; k < indptr[row + 1]; k += 32) {
        const int cidx = indices[k] - col_base;   // local column (may be < 0 or >= M in the halos)
        mn = min(mn, cidx);
        mx = max(mx, cidx);
        dg |= (cidx == row);
    }
    for (int o = 16; o > 0; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        dg |= __shfl_xor_sync(0xffffffffu, dg, o);
    }
    if (lane == 0) {
        lo[row] = mn;
        hi[row] = mx;
        has_diag[row] = dg;
    }
}

template <typename T>
__device__ __forceinline__ float reg_value(T v, bool diag, double s) {
    return (float)((1.0 - s) * (double)v + (diag ? s : 0.0));
}

__global__ void k_fill_f32(float* p, int64_t n, float v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

template <typename T>
__global__ void k_csr_to_dia(int64_t M, const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                             const T* __restrict__ data, float* __restrict__ band, int w, int64_t ldb, double s,
                             int col_base) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= M) return;
    for (int64_t k = indptr[row] + lane; k < indptr[row + 1]; k += 32) {
        const int cidx = indices[k] - col_base;
        const int d = cidx - (int)row + w;
        band[(int64_t)d * ldb + row] = reg_value(data[k], cidx == row, s);
    }
}

// symmetric half band: own rows give the upper diagonals, the couplings of own rows to the E extension
// rows before them (local column < 0) fill those rows' upper diagonals by symmetry
__global__ void k_dsym_fill_diag(float* U, int64_t M, int64_t E, int64_t ngr, float v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x)
        U[sgv_dsym_index(i + E, 0, ngr)] = v;
}

template <typename T>
__global__ void k_csr_to_dsym(int64_t M, const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                              const T* __restrict__ data, float* __restrict__ U, int64_t ngr, int64_t E, double s,
                              int col_base) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= M) return;
    for (int64_t k = indptr[row] + lane; k < indptr[row + 1]; k += 32) {
        const int64_t cl = (int64_t)indices[k] - col_base;
        if (cl > row) U[sgv_dsym_index(row + E, cl - row, ngr)] = reg_value(data[k], false, s);
        else if (cl == row) U[sgv_dsym_index(row + E, 0, ngr)] = 0.5f * reg_value(data[k], true, s);   // diagonal stored halved
        else if (cl < 0 && cl + E >= 0) U[sgv_dsym_index(cl + E, row - cl, ngr)] = reg_value(data[k], false, s);
    }
}

// entries below the diagonal inside the own rows must mirror the stored upper ones (to fp32 rounding)
template <typename T>
__global__ void k_dsym_check(int64_t M, const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                             const T* __restrict__ data, const float* __restrict__ U, int64_t ngr, int64_t E, double s,
                             int col_base, float abs_tol, unsigned long long* __restrict__ mismatches) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= M) return;
    unsigned bad = 0;
    long long delta = 0;   // stored non-zero entries above minus below the diagonal (own rows x own columns)
    for (int64_t k = indptr[row] + lane; k < indptr[row + 1]; k += 32) {
        const int64_t cl = (int64_t)indices[k] - col_base;
        if (cl > row && cl < M && data[k] != (T)0) ++delta;
        if (cl >= 0 && cl < row && data[k] != (T)0) --delta;
        if (cl >= 0 && cl < row) {
            // symmetric to fp32 rounding: relative to the entry, or (LD computed in fp32: the two triangles come
            // from different summation orders) to the largest diagonal entry
            const float lo = reg_value(data[k], false, s), up = U[sgv_dsym_index(cl + E, row - cl, ngr)];
            if (fabsf(lo - up) > fmaxf(4e-7f * fmaxf(fabsf(lo), fabsf(up)), abs_tol)) ++bad;
        }
    }
    if (bad) atomicAdd(mismatches, (unsigned long long)bad);
    // a lower entry is compared with its mirror above; an UPPER entry without a stored mirror (triu-only input) is
    // caught by the counts: all lower entries match and the counts are equal <=> the non-zero patterns mirror
    if (delta) atomicAdd(mismatches + 1, (unsigned long long)delta);
}

template <typename T>
__global__ void k_csr_to_panels(int64_t M, const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                const T* __restrict__ data, float* __restrict__ panels,
                                const int* __restrict__ blk_of_row, const int64_t* __restrict__ blk_start,
                                const int64_t* __restrict__ blk_off, const int* __restrict__ blk_ld, double s) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= M) return;
    const int b = blk_of_row[row];
    const int64_t s0 = blk_start[b];
    float* prow = panels + blk_off[b] + (row - s0) * blk_ld[b];
    for (int64_t k = indptr[row] + lane; k < indptr[row + 1]; k += 32) {
        const int cidx = indices[k];
        prow[cidx - s0] = reg_value(data[k], cidx == row, s);
    }
}

__global__ void k_panel_diag(int64_t M, float* __restrict__ panels, const int* __restrict__ blk_of_row,
                             const int64_t* __restrict__ blk_start, const int64_t* __restrict__ blk_off,
                             const int* __restrict__ blk_ld, float v) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= M) return;
    const int b = blk_of_row[row];
    const int64_t r = row - blk_start[b];
    panels[blk_off[b] + r * blk_ld[b] + r] = v;
}

template <typename T>
__global__ void k_csr_vals(int64_t M, const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                           const T* __restrict__ data, float* __restrict__ vals, double s, int col_base) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= M) return;
    for (int64_t k = indptr[row] + lane; k < indptr[row + 1]; k += 32)
        vals[k] = reg_value(data[k], indices[k] - col_base == row, s);
}

// dense chunk: src rows [r0, r0+nr) of an M-column row-major matrix with leading dimension ld_src
template <typename T>
__global__ void k_dense_convert(const T* __restrict__ src, int64_t ld_src, float* __restrict__ dst, int64_t ld_dst,
                                int64_t r0, int64_t nr, int64_t M, double s) {
    const int64_t n = nr * ld_dst;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / ld_dst, cidx = t - r * ld_dst;
        float v = 0.f;
        if (cidx < M) v = reg_value(src[r * ld_src + cidx], (r0 + r) == cidx, s);
        dst[(r0 + r) * ld_dst + cidx] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// symmetry of dense panels.  The default dense / block-diagonal kernel (spmm_psym.cu) reads only the upper triangle
// and the full-panel kernel computes P^T v, so both equal the reference's R @ v only for symmetric R.  Every panel is
// therefore compared with its transpose once, at upload (32 x 32 tile pairs through shared memory, one read of the
// panel); a panel store that is NOT symmetric is transposed in place and pinned to the full-panel kernel, whose
// P^T v is then exactly R v.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_panel_asym(const float* __restrict__ P, int m, int ld, float abs_tol, unsigned long long* __restrict__ bad) {
    const int bx = blockIdx.x, by = blockIdx.y;
    if (by > bx) return;
    __shared__ float tB[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int r = ty; r < 32; r += 8) {                         // B = P[bx*32 + r][by*32 + c]
        const int i = bx * 32 + r, j = by * 32 + tx;
        tB[r][tx] = (i < m && j < m) ? P[(int64_t)i * ld + j] : 0.f;
    }
    __syncthreads();
    unsigned n = 0;
    for (int r = ty; r < 32; r += 8) {                         // A = P[by*32 + r][bx*32 + c]  vs  B[c][r]
        const int i = by * 32 + r, j = bx * 32 + tx;
        if (i < m && j < m && j > i) {
            const float up = P[(int64_t)i * ld + j], lo = tB[tx][r];
            if (fabsf(lo - up) > fmaxf(4e-7f * fmaxf(fabsf(lo), fabsf(up)), abs_tol)) ++n;
        }
    }
    if (n) atomicAdd(bad, (unsigned long long)n);
}

__global__ void __launch_bounds__(256)
k_panel_transpose(float* __restrict__ P, int m, int ld) {
    const int bx = blockIdx.x, by = blockIdx.y;
    if (by > bx) return;
    __shared__ float tA[32][33], tB[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int ia = by * 32 + r, ja = bx * 32 + tx, ib = bx * 32 + r, jb = by * 32 + tx;
        tA[r][tx] = (ia < m && ja < m) ? P[(int64_t)ia * ld + ja] : 0.f;
        tB[r][tx] = (ib < m && jb < m) ? P[(int64_t)ib * ld + jb] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int ia = by * 32 + r, ja = bx * 32 + tx, ib = bx * 32 + r, jb = by * 32 + tx;
        if (ia < m && ja < m) P[(int64_t)ia * ld + ja] = tB[tx][r];
        if (bx != by && ib < m && jb < m) P[(int64_t)ib * ld + jb] = tA[tx][r];
    }
}

// returns 0 and sets ld.panel_sym; a non-symmetric store that the library does not own cannot be transposed: error
static int check_panel_symmetry(sgv_ctx* c, LdMatrix& ld, const std::vector<int64_t>& starts, const std::vector<int64_t>& offs,
                                const std::vector<int>& lds) {
    const int nb = (int)starts.size() - 1;
    unsigned long long* d_bad = reinterpret_cast<unsigned long long*>(c->counter + 8);
    SGV_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(unsigned long long), c->stream));
    float diag0 = 1.f;
    SGV_CUDA(cudaMemcpyAsync(&diag0, ld.panels + offs[0], sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    SGV_CUDA(cudaStreamSynchronize(c->stream));
    const float abs_tol = 2e-6f * fmaxf(fabsf(diag0), 1e-30f);
    for (int b = 0; b < nb; ++b) {
        const int m = (int)(starts[b + 1] - starts[b]);
        if (m < 2) continue;
        const unsigned nt = (unsigned)((m + 31) / 32);
        k_panel_asym<<<dim3(nt, nt), 256, 0, c->stream>>>(ld.panels + offs[b], m, lds[b], abs_tol, d_bad);
        c->launches++;
    }
    unsigned long long bad = 0;
    SGV_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost, c->stream));
    SGV_CUDA(cudaStreamSynchronize(c->stream));
    SGV_CUDA(cudaGetLastError());
    ld.panel_sym = bad == 0;
    if (bad == 0) return 0;
    SGV_CHECK(ld.owned, "adopted dense LD is not symmetric (%llu entries differ from their mirror): LD must be symmetric",
              bad);
    for (int b = 0; b < nb; ++b) {
        const int m = (int)(starts[b + 1] - starts[b]);
        if (m < 2) continue;
        const unsigned nt = (unsigned)((m + 31) / 32);
        k_panel_transpose<<<dim3(nt, nt), 256, 0, c->stream>>>(const_cast<float*>(ld.panels) + offs[b], m, lds[b]);
        c->launches++;
    }
    SGV_CUDA(cudaStreamSynchronize(c->stream));
    SGV_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// panel work items
// ---------------------------------------------------------------------------------------------
int sgv_build_panel_items(sgv_ctx* c, LdMatrix& ld, const std::vector<int64_t>& starts,
                          const std::vector<int64_t>& offs, const std::vector<int>& lds) {
    SGV_TRY(check_panel_symmetry(c, ld, starts, offs, lds));
    const int TI = 128 * ld.panel_rw;
    const int nb = (int)starts.size() - 1;
    int64_t tiles = 0;
    for (int b = 0; b < nb; ++b) tiles += (starts[b + 1] - starts[b] + TI - 1) / TI;
    const int64_t target = (int64_t)c->sm_count * 4;
    int s_cross = (int)std::min<int64_t>(64, std::max<int64_t>(1, (target + tiles - 1) / tiles));
    // do not cut segments shorter than 64 stored rows for the largest block
    int64_t mmax = 0;
    for (int b = 0; b < nb; ++b) mmax = std::max(mmax, starts[b + 1] - starts[b]);
    while (s_cross > 1 && mmax / s_cross < 64) --s_cross;
    std::vector<PanelItem> items;
    items.reserve((size_t)tiles * s_cross);
    for (int b = 0; b < nb; ++b) {
        const int64_t s0 = starts[b], m = starts[b + 1] - s0;
        const int64_t per = (m + s_cross - 1) / s_cross;
        for (int64_t i0 = 0; i0 < m; i0 += TI) {
            for (int sl = 0; sl < s_cross; ++sl) {
                PanelItem it;
                const int64_t j0 = std::min<int64_t>(m, (int64_t)sl * per);
                const int64_t nj = std::max<int64_t>(0, std::min<int64_t>(per, m - j0));
                it.ld = lds[b];
                it.off = offs[b] + j0 * lds[b] + i0;
                it.i0 = (int)(s0 + i0);
                it.ni = (int)std::min<int64_t>(TI, m - i0);
                it.j0 = (int)(s0 + j0);
                it.nj = (int)nj;
                it.navail = (int)std::min<int64_t>(TI, lds[b] - i0);
                it.slot = sl;
                items.push_back(it);
            }
        }
    }
    // heavier items first: better tail balance
    std::stable_sort(items.begin(), items.end(), [](const PanelItem& x, const PanelItem& y) {
        return (int64_t)x.ni * x.nj > (int64_t)y.ni * y.nj;
    });
    if (ld.items) cudaFree(ld.items);
    ld.items = nullptr;
    SGV_CUDA(cudaMalloc(&ld.items, items.size() * sizeof(PanelItem)));
    SGV_CUDA(cudaMemcpy(ld.items, items.data(), items.size() * sizeof(PanelItem), cudaMemcpyHostToDevice));
    ld.n_items = (int)items.size();
    ld.s_cross = s_cross;
    ld.nblocks = nb;
    const int64_t need = (int64_t)s_cross * c->Ml;
    if (c->ypart_cap < need) {
        SGV_CUDA(cudaStreamSynchronize(c->stream));
        if (c->ypart) cudaFree(c->ypart);
        c->ypart = nullptr;
        c->ypart_cap = 0;
        SGV_CUDA(cudaMalloc(&c->ypart, need * sizeof(double2)));
        c->ypart_cap = need;
    }
    return sgv_build_psym_items(c, ld, starts, offs, lds);   // + the upper-triangle items (default kernel)
}

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
static int check_cohort(sgv_ctx* c, int cohort) {
    SGV_CHECK(c != nullptr, "null handle");
    SGV_CHECK(c->M > 0, "sgv_configure has not been called");
    SGV_CHECK(cohort >= 0 && cohort < c->K, "cohort index %d out of range [0,%d)", cohort, c->K);
    SGV_CUDA(cudaSetDevice(c->device));
    return 0;
}

extern "C" int sgv_ld_upload_dense(sgv_handle c, int cohort, const void* R, int dtype, int64_t ld_src, double s) {
    SGV_TRY(check_cohort(c, cohort));
    SGV_CHECK(R != nullptr, "R is null");
    SGV_CHECK(dtype == SGV_F32 || dtype == SGV_F64, "bad dtype %d", dtype);
    SGV_CHECK(c->world == 1 && !c->rowpart, "whole-matrix dense upload is single-rank; a rows partition takes sgv_ld_upload_dense_rows");
    SGV_CHECK(ld_src >= c->M, "leading dimension %lld < M", (long long)ld_src);
    SGV_CUDA(cudaSetDevice(c->device));
    LdMatrix& ld = c->coh[cohort].ld;
    sgv_ld_free(ld);
    const int64_t M = c->M, ldd = round_up(M, 4);
    float* P = nullptr;
    SGV_CUDA(cudaMalloc(&P, (size_t)M * ldd * sizeof(float)));
    ld.panels = P;
    ld.owned = true;
    ld.layout = SGV_LAYOUT_DENSE;
    ld.nnz_stored = M * M;
    const size_t esz = dtype == SGV_F64 ? 8 : 4;
    const int64_t chunk_rows = std::max<int64_t>(1, std::min<int64_t>(M, (int64_t)(256 << 20) / (int64_t)(ld_src * esz)));
    SGV_TRY(sgv_ensure_stage(c, chunk_rows * ld_src * esz));
    for (int64_t r0 = 0; r0 < M; r0 += chunk_rows) {
        const int64_t nr = std::min(chunk_rows, M - r0);
        SGV_CUDA(cudaMemcpyAsync(c->stage, (const char*)R + (size_t)r0 * ld_src * esz, (size_t)nr * ld_src * esz,
                                 cudaMemcpyHostToDevice, c->stream));
        if (dtype == SGV_F64)
            k_dense_convert<double><<<1184, 256, 0, c->stream>>>((const double*)c->stage, ld_src, P, ldd, r0, nr, M, s);
        else
            k_dense_convert<float><<<1184, 256, 0, c->stream>>>((const float*)c->stage, ld_src, P, ldd, r0, nr, M, s);
        c->launches++;
        SGV_CUDA(cudaStreamSynchronize(c->stream));   // the staging buffer is reused
    }
    SGV_CUDA(cudaGetLastError());
    return sgv_build_panel_items(c, ld, {0, M}, {0}, {(int)ldd});
}

// ---------------------------------------------------------------------------------------------
// Dense LD partitioned by ROWS over the ranks (sgv_configure_part with halo = 2).  Rank r owns the markers
// [row_lo, row_hi) and holds the rows R[row_lo:row_hi, :] as the COLUMN PANEL  P[j][i] = R[row_lo + i][j]
// (M stored rows of Ml values): the full-panel kernel's column sweep  y[i] = sum_j P[j][i] v[j]  is then exactly
// (R v)[row_lo + i] - also for a non-symmetric R - with v the vector pair of ALL ranks (gathered from the peers'
// symmetric arenas by k_gather_rows, spmm.cu) and the outputs owned locally: no partial sums cross the ranks.
// ---------------------------------------------------------------------------------------------
// src: rows [r0, r0+nr) of this rank's slice (nr x M, row-major, leading dimension ld_src) -> dst[j][r0 + r]
template <typename T>
__global__ void __launch_bounds__(256)
k_rows_to_colpanel(const T* __restrict__ src, int64_t ld_src, float* __restrict__ dst, int64_t ld_dst, int64_t r0,
                   int64_t nr, int64_t M, int64_t row_lo, double s) {
    __shared__ float t[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t jt = (int64_t)blockIdx.x * 32, rt = (int64_t)blockIdx.y * 32;   // column tile of src, row tile of src
    for (int r = ty; r < 32; r += 8) {
        const int64_t rr = rt + r, j = jt + tx;
        t[r][tx] = (rr < nr && j < M) ? reg_value(src[rr * ld_src + j], (row_lo + r0 + rr) == j, s) : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int64_t j = jt + r, i = r0 + rt + tx;
        if (j < M && rt + tx < nr) dst[j * ld_dst + i] = t[tx][r];
    }
}

static int build_colpanel_items(sgv_ctx* c, LdMatrix& ld, int64_t ldd) {
    const int TI = 128 * ld.panel_rw;
    const int64_t M = c->M, Ml = c->Ml;
    const int64_t tiles = (Ml + TI - 1) / TI;
    const int64_t target = (int64_t)c->sm_count * 4;
    int s_cross = (int)std::min<int64_t>(64, std::max<int64_t>(1, (target + tiles - 1) / tiles));
    while (s_cross > 1 && M / s_cross < 64) --s_cross;
    const int64_t per = (M + s_cross - 1) / s_cross;
    std::vector<PanelItem> items;
    for (int64_t i0 = 0; i0 < Ml; i0 += TI) {
        for (int sl = 0; sl < s_cross; ++sl) {
            PanelItem it;
            const int64_t j0 = std::min<int64_t>(M, (int64_t)sl * per);
            it.ld = (int)ldd;
            it.off = j0 * ldd + i0;
            it.i0 = (int)i0;
            it.ni = (int)std::min<int64_t>(TI, Ml - i0);
            it.j0 = (int)j0;                       // GLOBAL marker index: the kernel reads the gathered vector
            it.nj = (int)std::max<int64_t>(0, std::min<int64_t>(per, M - j0));
            it.navail = (int)std::min<int64_t>(TI, ldd - i0);
            it.slot = sl;
            items.push_back(it);
        }
    }
    std::stable_sort(items.begin(), items.end(), [](const PanelItem& x, const PanelItem& y) {
        return (int64_t)x.ni * x.nj > (int64_t)y.ni * y.nj;
    });
    SGV_CUDA(cudaMalloc(&ld.items, items.size() * sizeof(PanelItem)));
    SGV_CUDA(cudaMemcpy(ld.items, items.data(), items.size() * sizeof(PanelItem), cudaMemcpyHostToDevice));
    ld.n_items = (int)items.size();
    ld.s_cross = s_cross;
    ld.nblocks = 1;
    ld.rowpart = true;
    ld.panel_sym = false;                          // never the upper-triangle kernel
    const int64_t need = (int64_t)s_cross * Ml;
    if (c->ypart_cap < need) {
        SGV_CUDA(cudaStreamSynchronize(c->stream));
        if (c->ypart) cudaFree(c->ypart);
        c->ypart = nullptr;
        c->ypart_cap = 0;
        SGV_CUDA(cudaMalloc(&c->ypart, need * sizeof(double2)));
        c->ypart_cap = need;
    }
    return 0;
}

extern "C" int sgv_ld_upload_dense_rows(sgv_handle c, int cohort, const void* rows, int dtype, int64_t ld_src, double s) {
    SGV_TRY(check_cohort(c, cohort));
    SGV_CHECK(rows != nullptr, "rows is null");
    SGV_CHECK(dtype == SGV_F32 || dtype == SGV_F64, "bad dtype %d", dtype);
    SGV_CHECK(c->rowpart, "the handle is not configured for the dense rows partition (sgv_configure_part halo = 2)");
    SGV_CHECK(ld_src >= c->M, "leading dimension %lld < M", (long long)ld_src);
    LdMatrix& ld = c->coh[cohort].ld;
    sgv_ld_free(ld);
    const int64_t M = c->M, Ml = c->Ml, ldd = round_up(Ml, 4);
    float* P = nullptr;
    SGV_CUDA(cudaMalloc(&P, (size_t)M * ldd * sizeof(float)));
    SGV_CUDA(cudaMemsetAsync(P, 0, (size_t)M * ldd * sizeof(float), c->stream));
    ld.panels = P;
    ld.owned = true;
    ld.layout = SGV_LAYOUT_DENSE;
    ld.nnz_stored = M * Ml;
    const size_t esz = dtype == SGV_F64 ? 8 : 4;
    const int64_t chunk_rows = std::max<int64_t>(1, std::min<int64_t>(Ml, (int64_t)(256 << 20) / (int64_t)(ld_src * esz)));
    SGV_TRY(sgv_ensure_stage(c, chunk_rows * ld_src * esz));
    for (int64_t r0 = 0; r0 < Ml; r0 += chunk_rows) {
        const int64_t nr = std::min(chunk_rows, Ml - r0);
        SGV_CUDA(cudaMemcpyAsync(c->stage, (const char*)rows + (size_t)r0 * ld_src * esz, (size_t)nr * ld_src * esz,
                                 cudaMemcpyHostToDevice, c->stream));
        const dim3 grid((unsigned)((M + 31) / 32), (unsigned)((nr + 31) / 32));
        if (dtype == SGV_F64)
            k_rows_to_colpanel<double><<<grid, 256, 0, c->stream>>>((const double*)c->stage, ld_src, P, ldd, r0, nr, M, c->row_lo, s);
        else
            k_rows_to_colpanel<float><<<grid, 256, 0, c->stream>>>((const float*)c->stage, ld_src, P, ldd, r0, nr, M, c->row_lo, s);
        c->launches++;
        SGV_CUDA(cudaStreamSynchronize(c->stream));   // the staging buffer is reused
    }
    SGV_CUDA(cudaGetLastError());
    return build_colpanel_items(c, ld, ldd);
}

extern "C" int sgv_ld_adopt_dense_colpanel(sgv_handle c, int cohort, const float* P_dev, int64_t ldd) {
    SGV_TRY(check_cohort(c, cohort));
    SGV_CHECK(P_dev != nullptr && ((uintptr_t)P_dev & 15) == 0, "device pointer must be 16-byte aligned");
    SGV_CHECK(c->rowpart, "the handle is not configured for the dense rows partition (sgv_configure_part halo = 2)");
    SGV_CHECK(ldd >= c->Ml && ldd % 4 == 0, "ld must be >= the local row count and a multiple of 4");
    LdMatrix& ld = c->coh[cohort].ld;
    sgv_ld_free(ld);
    ld.panels = P_dev;
    ld.owned = false;
    ld.layout = SGV_LAYOUT_DENSE;
    ld.nnz_stored = c->M * c->Ml;
    return build_colpanel_items(c, ld, ldd);
}

extern "C" int sgv_ld_adopt_dense(sgv_handle c, int cohort, const float* R_dev, int64_t ldd) {
    SGV_TRY(check_cohort(c, cohort));
    SGV_CHECK(R_dev != nullptr && ((uintptr_t)R_dev & 15) == 0, "device pointer must be 16-byte aligned");
    SGV_CHECK(c->world == 1 && !c->rowpart, "whole-matrix dense LD is single-rank; a rows partition takes sgv_ld_adopt_dense_colpanel");
    SGV_CHECK(ldd >= c->M && ldd % 4 == 0, "ld must be >= M and a multiple of 4");
    LdMatrix& ld = c->coh[cohort].ld;
    sgv_ld_free(ld);
    ld.panels = R_dev;
    ld.owned = false;
    ld.layout = SGV_LAYOUT_DENSE;
    ld.nnz_stored = c->M * c->M;
    return sgv_build_panel_items(c, ld, {0, c->M}, {0}, {(int)ldd});
}

extern "C" int sgv_ld_adopt_blockdiag(sgv_handle c, int cohort, const float* panels_dev, int nblocks, const int64_t* starts,
                                      const int64_t* offs, const int* lds) {
    SGV_TRY(check_cohort(c, cohort));
    SGV_CHECK(panels_dev != nullptr && ((uintptr_t)panels_dev & 15) == 0, "device pointer must be 16-byte aligned");
    SGV_CHECK(nblocks >= 1 && starts && offs && lds, "null / empty block description");
    SGV_CHECK(starts[0] == 0 && starts[nblocks] == c->Ml, "blocks must cover the local rows [0,%lld)", (long long)c->Ml);
    SGV_CHECK(c->world == 1 || !c->halo, "block-diagonal LD is sharded at block boundaries (halo = 0)");
    std::vector<int64_t> st(starts, starts + nblocks + 1), of(offs, offs + nblocks);
    std::vector<int> ld_(lds, lds + nblocks);
    LdMatrix& ld = c->coh[cohort].ld;
    sgv_ld_free(ld);
    ld.nnz_stored = 0;
    for (int b = 0; b < nblocks; ++b) {
        const int64_t m = st[b + 1] - st[b];
        SGV_CHECK(m >= 1 && ld_[b] >= m && ld_[b] % 4 == 0 && of[b] % 4 == 0, "block %d: ld must be >= size and a multiple of 4, "
                  "offset a multiple of 4", b);
        ld.nnz_stored += m * m;
    }
    ld.panels = panels_dev;
    ld.owned = false;
    ld.layout = nblocks == 1 ? SGV_LAYOUT_DENSE : SGV_LAYOUT_BLOCKDIAG;
    return sgv_build_panel_items(c, ld, st, of, ld_);
}

extern "C" int sgv_ld_adopt_dia(sgv_handle c, int cohort, const float* band_dev, int64_t w, int64_t ldb) {
    SGV_TRY(check_cohort(c, cohort));
    SGV_CHECK(band_dev != nullptr && ((uintptr_t)band_dev & 15) == 0, "device pointer must be 16-byte aligned");
    SGV_CHECK(ldb >= c->Ml && ldb % 4 == 0, "ldb must be >= the local row count and a multiple of 4");
    SGV_CHECK(w >= 0 && sgv_dia_feasible(w), "half-bandwidth %lld not supported by the DIA kernel", (long long)w);
    LdMatrix& ld = c->coh[cohort].ld;
    sgv_ld_free(ld);
    ld.band = band_dev;
    ld.owned = false;
    ld.layout = SGV_LAYOUT_DIA;
    ld.w = w;
    ld.ldb = ldb;
    ld.nnz_stored = (2 * w + 1) * c->Ml;
    return 0;
}

extern "C" int sgv_dsym_extension(sgv_handle c, int64_t w, int64_t* ext) {
    SGV_CHECK(c != nullptr && ext != nullptr, "null argument");
    *ext = sgv_dsym_ext(c, w);
    return 0;
}

extern "C" int sgv_ld_adopt_dsym(sgv_handle c, int cohort, const float* U_dev, int64_t w, int64_t ldb, int64_t ext) {
    SGV_TRY(check_cohort(c, cohort));
    SGV_CHECK(U_dev != nullptr && ((uintptr_t)U_dev & 15) == 0, "device pointer must be 16-byte aligned");
    SGV_CHECK(w >= 0 && sgv_dsym_feasible(w), "half-bandwidth %lld not supported by the DSYM kernel", (long long)w);
    SGV_CHECK(ext == sgv_dsym_ext(c, w), "extension rows %lld, expected %lld (sgv_dsym_extension)", (long long)ext,
              (long long)sgv_dsym_ext(c, w));
    SGV_CHECK(ldb >= c->Ml + ext && ldb % 128 == 0, "ldb must be a multiple of 128 and >= local rows + extension");
    LdMatrix& ld = c->coh[cohort].ld;
    sgv_ld_free(ld);
    ld.band = U_dev;
    ld.owned = false;
    ld.layout = SGV_LAYOUT_DSYM;
    ld.w = w;
    ld.ldb = ldb;
    ld.ext = ext;
    ld.nnz_stored = (w + 1) * c->Ml;
    return sgv_dsym_ensure_scratch(c, ld);
}

template <typename T>
static int convert_csr(sgv_ctx* c, LdMatrix& ld, int layout, const int64_t* d_indptr, const int32_t* d_indices,
                       const T* d_data, int64_t nnz, double s, int64_t w, const std::vector<int64_t>& starts) {
    const int64_t M = c->Ml;
    const int col_base = c->halo ? (int)c->row_lo : 0;
    const unsigned wgrid = (unsigned)((M * 32 + 255) / 256);
    if (layout == SGV_LAYOUT_DSYM) {
        const int64_t E = sgv_dsym_ext(c, w), Dp = round_up(w + 1, 4);
        const int64_t ldb = round_up(M + E, 128), ngr = Dp / 4;
        float* U = nullptr;
        SGV_CUDA(cudaMalloc(&U, (size_t)Dp * ldb * sizeof(float)));
        ld.band = U;
        ld.owned = true;
        ld.w = w;
        ld.ldb = ldb;
        ld.ext = E;
        ld.nnz_stored = (w + 1) * M;
        SGV_CUDA(cudaMemsetAsync(U, 0, (size_t)Dp * ldb * sizeof(float), c->stream));
        k_dsym_fill_diag<<<592, 256, 0, c->stream>>>(U, M, E, ngr, 0.5f * (float)s);   // absent diagonal entry (stored halved)
        k_csr_to_dsym<T><<<wgrid, 256, 0, c->stream>>>(M, d_indptr, d_indices, d_data, U, ngr, E, s, col_base);
        unsigned long long* d_bad = reinterpret_cast<unsigned long long*>(c->counter + 8);   // spare words of the ticket block
        SGV_CUDA(cudaMemsetAsync(d_bad, 0, 2 * sizeof(unsigned long long), c->stream));
        // absolute tolerance 2e-6 x the scale of the diagonal (1 for a correlation matrix; read from row 0)
        float diag0 = 1.f;
        SGV_CUDA(cudaMemcpyAsync(&diag0, U + sgv_dsym_index(E, 0, ngr), sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        SGV_CUDA(cudaStreamSynchronize(c->stream));
        const float abs_tol = 2e-6f * fmaxf(2.f * fabsf(diag0), 1e-30f);
        k_dsym_check<T><<<wgrid, 256, 0, c->stream>>>(M, d_indptr, d_indices, d_data, U, ngr, E, s, col_base, abs_tol, d_bad);
        c->launches += 3;
        unsigned long long bad2[2] = {0, 0};
        SGV_CUDA(cudaMemcpyAsync(bad2, d_bad, sizeof(bad2), cudaMemcpyDeviceToHost, c->stream));
        SGV_CUDA(cudaStreamSynchronize(c->stream));
        if (bad2[0] != 0 || bad2[1] != 0) {   // values or pattern not symmetric: the caller falls back to the full band
            sgv_ld_free(ld);
            return 1;
        }
        ld.layout = layout;
        SGV_TRY(sgv_dsym_ensure_scratch(c, ld));
    } else if (layout == SGV_LAYOUT_DIA) {
        const int64_t ldb = round_up(M, 32);
        float* band = nullptr;
        SGV_CUDA(cudaMalloc(&band, (size_t)(2 * w + 1) * ldb * sizeof(float)));
        ld.band = band;
        ld.owned = true;
        ld.w = w;
        ld.ldb = ldb;
        ld.nnz_stored = (2 * w + 1) * M;
        SGV_CUDA(cudaMemsetAsync(band, 0, (size_t)(2 * w + 1) * ldb * sizeof(float), c->stream));
        k_fill_f32<<<592, 256, 0, c->stream>>>(band + w * ldb, M, (float)s);   // value of an absent diagonal entry
        k_csr_to_dia<T><<<wgrid, 256, 0, c->stream>>>(M, d_indptr, d_indices, d_data, band, (int)w, ldb, s, col_base);
        c->launches += 2;
    } else if (layout == SGV_LAYOUT_DENSE || layout == SGV_LAYOUT_BLOCKDIAG) {
        const int nb = (int)starts.size() - 1;
        std::vector<int64_t> offs(nb);
        std::vector<int> lds(nb), blk_of_row(M);
        int64_t total = 0;
        for (int b = 0; b < nb; ++b) {
            const int64_t m = starts[b + 1] - starts[b];
            lds[b] = (int)round_up(m, 4);
            offs[b] = total;
            total += m * lds[b];
            for (int64_t i = starts[b]; i < starts[b + 1]; ++i) blk_of_row[i] = b;
        }
        float* P = nullptr;
        SGV_CUDA(cudaMalloc(&P, (size_t)total * sizeof(float)));
        ld.panels = P;
        ld.owned = true;
        ld.nnz_stored = 0;
        for (int b = 0; b < nb; ++b) ld.nnz_stored += (starts[b + 1] - starts[b]) * (starts[b + 1] - starts[b]);
        int *d_bor = nullptr, *d_ld = nullptr;
        int64_t *d_start = nullptr, *d_off = nullptr;
        SGV_CUDA(cudaMalloc(&d_bor, M * sizeof(int)));
        SGV_CUDA(cudaMalloc(&d_ld, nb * sizeof(int)));
        SGV_CUDA(cudaMalloc(&d_start, (nb + 1) * sizeof(int64_t)));
        SGV_CUDA(cudaMalloc(&d_off, nb * sizeof(int64_t)));
        SGV_CUDA(cudaMemcpyAsync(d_bor, blk_of_row.data(), M * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        SGV_CUDA(cudaMemcpyAsync(d_ld, lds.data(), nb * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        SGV_CUDA(cudaMemcpyAsync(d_start, starts.data(), (nb + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
        SGV_CUDA(cudaMemcpyAsync(d_off, offs.data(), nb * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
        SGV_CUDA(cudaMemsetAsync(P, 0, (size_t)total * sizeof(float), c->stream));
        k_panel_diag<<<(unsigned)((M + 255) / 256), 256, 0, c->stream>>>(M, P, d_bor, d_start, d_off, d_ld, (float)s);
        k_csr_to_panels<T><<<wgrid, 256, 0, c->stream>>>(M, d_indptr, d_indices, d_data, P, d_bor, d_start, d_off, d_ld, s);
        c->launches += 2;
        SGV_CUDA(cudaStreamSynchronize(c->stream));
        cudaFree(d_bor);
        cudaFree(d_ld);
        cudaFree(d_start);
        cudaFree(d_off);
        SGV_TRY(sgv_build_panel_items(c, ld, starts, offs, lds));
    } else {
        float* vals = nullptr;
        SGV_CUDA(cudaMalloc(&vals, (size_t)std::max<int64_t>(nnz, 1) * sizeof(float)));
        ld.vals = vals;
        ld.nnz = nnz;
        ld.nnz_stored = nnz;
        k_csr_vals<T><<<wgrid, 256, 0, c->stream>>>(M, d_indptr, d_indices, d_data, vals, s, c->rowpart ? (int)c->row_lo : 0);
        ld.rowpart = c->rowpart;     // rows partition: global column indices, the product reads the gathered vector pair
        c->launches++;
    }
    SGV_CUDA(cudaStreamSynchronize(c->stream));
    SGV_CUDA(cudaGetLastError());
    ld.layout = layout;
    return 0;
}

// SGV_TIMING=1: print the phases of an upload (host clock after a stream synchronise) to stderr
struct PhaseTimer {
    bool on;
    cudaStream_t st;
    std::chrono::steady_clock::time_point t;
    PhaseTimer(cudaStream_t s) : on(getenv("SGV_TIMING") != nullptr), st(s), t(std::chrono::steady_clock::now()) {}
    void lap(const char* what) {
        if (!on) return;
        cudaStreamSynchronize(st);
        const auto n = std::chrono::steady_clock::now();
        fprintf(stderr, "[sgv upload] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};

// device staging buffer freed on every exit path unless released to a longer-lived owner
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    void* release() {
        void* q = p;
        p = nullptr;
        return q;
    }
};

extern "C" int sgv_ld_upload_csr(sgv_handle c, int cohort, const int64_t* indptr, const int32_t* indices,
                                 const void* data, int dtype, int64_t nnz, double s, int layout_hint) {
    SGV_TRY(check_cohort(c, cohort));
    PhaseTimer pt(c->stream);
    SGV_CHECK(indptr && (nnz == 0 || (indices && data)), "null CSR arrays");
    SGV_CHECK(dtype == SGV_F32 || dtype == SGV_F64, "bad dtype %d", dtype);
    SGV_CHECK(c->M < INT_MAX, "M too large for int32 column indices");
    SGV_CUDA(cudaSetDevice(c->device));
    const int64_t M = c->Ml;   // local rows; columns are global when the partition has halos / is the rows partition
    const int col_base = (c->halo || c->rowpart) ? (int)c->row_lo : 0;
    SGV_CHECK(!c->rowpart || layout_hint == SGV_LAYOUT_AUTO || layout_hint == SGV_LAYOUT_CSR,
              "the rows partition keeps sparse LD in CSR form (dense LD: sgv_ld_upload_dense_rows)");
    SGV_CHECK(indptr[0] == 0 && indptr[M] == nnz, "indptr[0]=%lld indptr[M]=%lld inconsistent with nnz=%lld",
              (long long)indptr[0], (long long)indptr[M], (long long)nnz);
    LdMatrix& ld = c->coh[cohort].ld;
    sgv_ld_free(ld);
    const size_t esz = dtype == SGV_F64 ? 8 : 4;
    DevBuf b_indptr, b_indices, b_data, b_lo;   // freed on every return path (error paths included)
    SGV_CUDA(cudaMalloc(&b_indptr.p, (M + 1) * sizeof(int64_t)));
    SGV_CUDA(cudaMalloc(&b_indices.p, std::max<int64_t>(nnz, 1) * sizeof(int32_t)));
    SGV_CUDA(cudaMalloc(&b_data.p, std::max<int64_t>(nnz, 1) * esz));
    SGV_CUDA(cudaMalloc(&b_lo.p, 3 * M * sizeof(int)));
    int64_t* d_indptr = static_cast<int64_t*>(b_indptr.p);
    int32_t* d_indices = static_cast<int32_t*>(b_indices.p);
    void* d_data = b_data.p;
    int *d_lo = static_cast<int*>(b_lo.p), *d_hi = d_lo + M, *d_dg = d_hi + M;
    pt.lap("free old + cudaMalloc staging");
    SGV_CUDA(cudaMemcpyAsync(d_indptr, indptr, (M + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    SGV_CUDA(cudaMemcpyAsync(d_indices, indices, nnz * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    SGV_CUDA(cudaMemcpyAsync(d_data, data, nnz * esz, cudaMemcpyHostToDevice, c->stream));
    pt.lap("H2D copies");
    const unsigned wgrid = (unsigned)((M * 32 + 255) / 256);
    k_row_extent<<<wgrid, 256, 0, c->stream>>>(M, d_indptr, d_indices, d_lo, d_hi, d_dg, col_base);
    c->launches++;
    std::vector<int> ext(3 * M);
    SGV_CUDA(cudaMemcpyAsync(ext.data(), d_lo, 3 * M * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    SGV_CUDA(cudaStreamSynchronize(c->stream));
    const int *lo = ext.data(), *hi = lo + M, *dg = hi + M;

    pt.lap("row extents");
    // ---- structure analysis on M-length arrays ----
    int64_t w = 0;
    bool all_diag = true;
    for (int64_t i = 0; i < M; ++i) {
        if (hi[i] < 0) { all_diag = false; continue; }
        SGV_CHECK(lo[i] + col_base >= 0 && hi[i] + col_base < c->M, "column index out of range in row %lld", (long long)i);
        SGV_CHECK(c->halo || c->rowpart || (lo[i] >= 0 && hi[i] < M), "row %lld has columns outside this rank's shard", (long long)i);
        w = std::max<int64_t>(w, std::max<int64_t>(i - lo[i], hi[i] - i));
        all_diag = all_diag && dg[i];
    }
    std::vector<int> sufmin(M + 1);
    sufmin[M] = INT_MAX;
    for (int64_t i = M - 1; i >= 0; --i) sufmin[i] = std::min(sufmin[i + 1], lo[i]);
    std::vector<int64_t> starts;
    starts.push_back(0);
    int runmax = -1;
    for (int64_t i = 0; i < M; ++i) {
        if (i > 0 && runmax < i && sufmin[i] >= i) starts.push_back(i);
        runmax = std::max(runmax, hi[i]);
    }
    starts.push_back(M);
    const int64_t nb = (int64_t)starts.size() - 1;
    double dense_cells = 0;
    for (int64_t b = 0; b < nb; ++b) dense_cells += (double)(starts[b + 1] - starts[b]) * (double)(starts[b + 1] - starts[b]);
    const double dia_cells = (double)M * (2 * w + 1) - (double)w * (w + 1);
    const double fill_blk = nnz / std::max(dense_cells, 1.0), fill_dia = nnz / std::max(dia_cells, 1.0);
    size_t free_b = 0, total_b = 0;
    SGV_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const bool blk_fits = dense_cells * 4.0 < 0.8 * (double)free_b;
    const bool dia_ok = sgv_dia_feasible(w) && dia_cells * 4.0 < 0.8 * (double)free_b;

    if (c->bandwidth_hint > w) w = c->bandwidth_hint;
    int layout = layout_hint;
    const bool dsym_ok = sgv_dsym_feasible(w) && dia_cells * 2.0 < 0.8 * (double)free_b;
    bool dsym_fallback = false;   // DSYM chosen automatically: fall back to the full band if R is not symmetric
    if (c->world > 1 && c->halo) {
        SGV_CHECK(layout == SGV_LAYOUT_AUTO || layout == SGV_LAYOUT_DIA || layout == SGV_LAYOUT_DSYM,
                  "a row partition with halos needs a band layout (dia / dsym)");
        if (layout == SGV_LAYOUT_AUTO) {
            layout = dsym_ok ? SGV_LAYOUT_DSYM : SGV_LAYOUT_DIA;
            dsym_fallback = dsym_ok;
        }
    }
    if (c->rowpart) layout = SGV_LAYOUT_CSR;      // rows of a general sparse R: any column of the matrix, indices stay global
    if (layout == SGV_LAYOUT_AUTO) {
        layout = SGV_LAYOUT_CSR;
        const bool blk_good = blk_fits && fill_blk >= (nb == 1 ? 0.25 : 0.5);
        const bool dia_good = dia_ok && fill_dia >= 0.35;
        if (blk_good && dia_good) layout = (dense_cells <= dia_cells) ? SGV_LAYOUT_BLOCKDIAG : SGV_LAYOUT_DIA;
        else if (blk_good) layout = SGV_LAYOUT_BLOCKDIAG;
        else if (dia_good) layout = SGV_LAYOUT_DIA;
        if (layout == SGV_LAYOUT_BLOCKDIAG && nb == 1) layout = SGV_LAYOUT_DENSE;
        if (layout == SGV_LAYOUT_DIA && dsym_ok) {
            layout = SGV_LAYOUT_DSYM;
            dsym_fallback = true;
        }
    }
    int rc = 0;
    if (layout == SGV_LAYOUT_DSYM && !dsym_ok) { sgv_set_error("DSYM layout infeasible for half-bandwidth %lld", (long long)w); rc = -1; }
    if (layout == SGV_LAYOUT_DIA && !dia_ok) { sgv_set_error("DIA layout infeasible for half-bandwidth %lld", (long long)w); rc = -1; }
    if ((layout == SGV_LAYOUT_DENSE || layout == SGV_LAYOUT_BLOCKDIAG) && !blk_fits) { sgv_set_error("dense blocks do not fit in device memory"); rc = -1; }
    if (layout == SGV_LAYOUT_CSR && s != 0.0 && !all_diag) { sgv_set_error("CSR layout with s != 0 needs every diagonal entry stored"); rc = -3; }
    if (layout == SGV_LAYOUT_DENSE) starts = {0, M};
    pt.lap("structure analysis (host)");
    for (int attempt = 0; rc == 0 && attempt < 2; ++attempt) {
        if (dtype == SGV_F64) rc = convert_csr<double>(c, ld, layout, d_indptr, d_indices, (const double*)d_data, nnz, s, w, starts);
        else rc = convert_csr<float>(c, ld, layout, d_indptr, d_indices, (const float*)d_data, nnz, s, w, starts);
        if (rc != 1) break;                       // rc == 1: the DSYM conversion found R not symmetric
        if (dsym_fallback && dia_ok) {
            layout = SGV_LAYOUT_DIA;
            rc = 0;
        } else {
            sgv_set_error("LD matrix is not symmetric: the DSYM layout cannot hold it");
            rc = -1;
        }
    }
    pt.lap("layout conversion");
    if (rc == 0 && layout == SGV_LAYOUT_CSR) {   // the CSR layout keeps the index arrays
        ld.indptr = static_cast<int64_t*>(b_indptr.release());
        ld.indices = static_cast<int32_t*>(b_indices.release());
    }
    if (rc != 0) sgv_ld_free(ld);
    pt.lap("free staging");
    return rc;
}

// ---------------------------------------------------------------------------------------------
// LD given in scipy's DIA format (what banded LD is naturally stored as): no index arrays to ship,
// and for the symmetric half-band layout only the upper diagonals are needed on the device.
// ---------------------------------------------------------------------------------------------
// stage[kk*n + t] = value of diagonal `off` at global row g0 + t  (= R[g0+t][g0+t+off]), kk-th diagonal of the chunk
template <typename T>
__global__ void k_dia_to_dsym(const T* __restrict__ stage, int64_t n, const int* __restrict__ offs, int nk, float* __restrict__ U,
                              int64_t ngr, int64_t g0, int64_t row_lo, int64_t M, double s) {
    const int kk = blockIdx.y;
    if (kk >= nk) return;
    const int off = offs[kk];
    const T* src = stage + (int64_t)kk * n;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t g = g0 + t;                       // global row; storage row = t (the buffer starts at g0 = row_lo - E)
        float v = 0.f;
        if (g >= 0 && g + off < M && g + off >= row_lo)  // extension rows keep only their couplings to own rows
            v = off == 0 ? 0.5f * reg_value(src[t], true, s) : reg_value(src[t], false, s);
        U[sgv_dsym_index(t, off, ngr)] = v;
    }
}

// lower diagonals against the stored upper ones: R[g][g-off] (lower, off > 0) must equal R[g-off][g]
template <typename T>
__global__ void k_dia_check_lower(const T* __restrict__ stage, int64_t n, const int* __restrict__ offs, int nk,
                                  const float* __restrict__ U, int64_t ngr, int64_t g0, int64_t row_lo, int64_t M, double s,
                                  float abs_tol, unsigned long long* __restrict__ mismatches) {
    const int kk = blockIdx.y;
    if (kk >= nk) return;
    const int off = offs[kk];                           // > 0: the diagonal -off
    const T* src = stage + (int64_t)kk * n;
    unsigned bad = 0;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t g = g0 + t;                       // row of the lower entry; its mirror is row g - off
        if (g >= row_lo && g < M && g - off >= g0 && g - off >= 0) {
            const float lo = reg_value(src[t], false, s), up = U[sgv_dsym_index(t - off, off, ngr)];
            if (fabsf(lo - up) > fmaxf(4e-7f * fmaxf(fabsf(lo), fabsf(up)), abs_tol)) ++bad;
        }
    }
    if (bad) atomicAdd(mismatches, (unsigned long long)bad);
}

template <typename T>
__global__ void k_dia_to_band(const T* __restrict__ stage, int64_t n, const int* __restrict__ offs, int nk, float* __restrict__ band,
                              int64_t w, int64_t ldb, int64_t g0, int64_t M, double s) {
    const int kk = blockIdx.y;
    if (kk >= nk) return;
    const int off = offs[kk];
    const T* src = stage + (int64_t)kk * n;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t g = g0 + t;
        float v = 0.f;
        if (g + off >= 0 && g + off < M) v = reg_value(src[t], off == 0, s);
        band[(int64_t)(off + w) * ldb + t] = v;
    }
}

// Copies, for the diagonals listed in ks, the entries of global rows [g0, g0+n) into the staging buffer (chunks of
// diagonals; host -> device on the copy stream, double-buffered against the conversion kernel on the compute stream)
// and calls `convert(stage_ptr, dev_offsets, nk)` for every chunk.
template <typename F>
static int stream_diagonals(sgv_ctx* c, const void* data, size_t esz, int64_t ldd, int64_t col0, const int64_t* offsets,
                            const std::vector<int>& ks, bool negate, int64_t g0, int64_t n, int64_t M, F convert) {
    const int CH = 16;
    const size_t chunk_bytes = (size_t)CH * n * esz;
    SGV_TRY(sgv_ensure_stage(c, 2 * (int64_t)chunk_bytes + 2 * CH * (int64_t)sizeof(int)));
    char* base = static_cast<char*>(c->stage);
    int* d_offs = reinterpret_cast<int*>(base + 2 * chunk_bytes);
    cudaEvent_t h2d[2], done[2];
    for (int i = 0; i < 2; ++i) {
        SGV_CUDA(cudaEventCreateWithFlags(&h2d[i], cudaEventDisableTiming));
        SGV_CUDA(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
    }
    SGV_CUDA(cudaStreamSynchronize(c->stream));
    int rc = 0;
    std::vector<int> hoffs(2 * CH);
    for (size_t c0 = 0, ci = 0; c0 < ks.size() && rc == 0; c0 += CH, ++ci) {
        const int b = (int)(ci & 1), nk = (int)std::min<size_t>(CH, ks.size() - c0);
        char* st = base + (size_t)b * chunk_bytes;
        if (ci >= 2) SGV_CUDA(cudaStreamWaitEvent(c->copy_stream, done[b], 0));
        SGV_CUDA(cudaMemsetAsync(st, 0, chunk_bytes, c->copy_stream));
        for (int kk = 0; kk < nk; ++kk) {
            const int k = ks[c0 + kk];
            const int64_t off = offsets[k];
            hoffs[b * CH + kk] = (int)(negate ? -off : off);
            // scipy DIA: the entry of diagonal `off` in row g sits at column j = g + off of data row k
            int64_t j_lo = g0 + off, j_hi = g0 + n + off;                    // wanted columns [j_lo, j_hi)
            const int64_t a_lo = std::max<int64_t>(std::max<int64_t>(j_lo, col0), 0);
            const int64_t a_hi = std::min<int64_t>(std::min<int64_t>(j_hi, col0 + ldd), M);
            if (a_hi > a_lo)
                SGV_CUDA(cudaMemcpyAsync(st + ((size_t)kk * n + (size_t)(a_lo - j_lo)) * esz,
                                         static_cast<const char*>(data) + ((size_t)k * ldd + (size_t)(a_lo - col0)) * esz,
                                         (size_t)(a_hi - a_lo) * esz, cudaMemcpyHostToDevice, c->copy_stream));
        }
        SGV_CUDA(cudaMemcpyAsync(d_offs + b * CH, hoffs.data() + b * CH, nk * sizeof(int), cudaMemcpyHostToDevice, c->copy_stream));
        SGV_CUDA(cudaEventRecord(h2d[b], c->copy_stream));
        SGV_CUDA(cudaStreamWaitEvent(c->stream, h2d[b], 0));
        rc = convert(st, d_offs + b * CH, nk);
        c->launches++;
        SGV_CUDA(cudaEventRecord(done[b], c->stream));
    }
    SGV_CUDA(cudaStreamSynchronize(c->stream));
    SGV_CUDA(cudaStreamSynchronize(c->copy_stream));
    for (int i = 0; i < 2; ++i) {
        cudaEventDestroy(h2d[i]);
        cudaEventDestroy(done[i]);
    }
    SGV_CUDA(cudaGetLastError());
    return rc;
}

template <typename T>
static int upload_dia_t(sgv_ctx* c, LdMatrix& ld, const T* data, int64_t ldd, int64_t col0, const int64_t* offsets, int ndiag,
                        double s, int layout, int assume_symmetric, int64_t w, PhaseTimer& pt) {
    const int64_t M = c->M, Ml = c->Ml;
    std::vector<int> up, lowr, all;
    for (int k = 0; k < ndiag; ++k) {
        all.push_back(k);
        if (offsets[k] >= 0) up.push_back(k);
        else lowr.push_back(k);
    }
    const dim3 blk(256);
    if (layout == SGV_LAYOUT_DSYM && !assume_symmetric) {
        // an upper diagonal whose mirror is not among the containers' diagonals: the matrix is not symmetric (only
        // stored diagonals can be compared below), the half band cannot hold it
        for (int ku : up) {
            if (offsets[ku] == 0) continue;
            bool found = false;
            for (int kl : lowr) found = found || offsets[kl] == -offsets[ku];
            if (!found) return 1;
        }
    }
    if (layout == SGV_LAYOUT_DSYM) {
        const int64_t E = sgv_dsym_ext(c, w), Dp = round_up(w + 1, 4), ngr = Dp / 4;
        const int64_t n = round_up(Ml + E, 128), g0 = c->row_lo - E;
        float* U = nullptr;
        SGV_CUDA(cudaMalloc(&U, (size_t)Dp * n * sizeof(float)));
        ld.band = U;
        ld.owned = true;
        ld.w = w;
        ld.ldb = n;
        ld.ext = E;
        ld.nnz_stored = (w + 1) * Ml;
        pt.lap("cudaMalloc half band");
        SGV_CUDA(cudaMemsetAsync(U, 0, (size_t)Dp * n * sizeof(float), c->stream));
        pt.lap("zero fill");
        const int64_t row_lo = c->row_lo;
        // rows of the buffer beyond the local range (padding to 128) convert to zeros: g + off < M fails or data is zero
        k_dsym_fill_diag<<<592, 256, 0, c->stream>>>(U, Ml, E, ngr, 0.5f * (float)s);
        const int64_t n_rows = std::min<int64_t>(n, Ml + E);
        auto conv = [&](char* st, int* d_offs, int nk) -> int {
            const dim3 grid((unsigned)std::min<int64_t>((n_rows + 255) / 256, 2048), (unsigned)nk);
            k_dia_to_dsym<T><<<grid, blk, 0, c->stream>>>((const T*)st, n_rows, d_offs, nk, U, ngr, g0, row_lo, M, s);
            return 0;
        };
        SGV_TRY(stream_diagonals(c, data, sizeof(T), ldd, col0, offsets, up, false, g0, n_rows, M, conv));
        pt.lap("upper diagonals H2D + convert");
        if (!assume_symmetric && !lowr.empty()) {
            unsigned long long* d_bad = reinterpret_cast<unsigned long long*>(c->counter + 8);
            SGV_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(unsigned long long), c->stream));
            float diag0 = 1.f;
            SGV_CUDA(cudaMemcpyAsync(&diag0, U + sgv_dsym_index(E, 0, ngr), sizeof(float), cudaMemcpyDeviceToHost, c->stream));
            SGV_CUDA(cudaStreamSynchronize(c->stream));
            const float abs_tol = 2e-6f * fmaxf(2.f * fabsf(diag0), 1e-30f);
            auto chk = [&](char* st, int* d_offs, int nk) -> int {
                const dim3 grid((unsigned)std::min<int64_t>((n_rows + 255) / 256, 2048), (unsigned)nk);
                k_dia_check_lower<T><<<grid, blk, 0, c->stream>>>((const T*)st, n_rows, d_offs, nk, U, ngr, g0, row_lo, M, s, abs_tol,
                                                                  d_bad);
                return 0;
            };
            SGV_TRY(stream_diagonals(c, data, sizeof(T), ldd, col0, offsets, lowr, true, g0, n_rows, M, chk));
            unsigned long long bad = 0;
            SGV_CUDA(cudaMemcpy(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost));
            if (bad != 0) {
                sgv_ld_free(ld);
                return 1;   // not symmetric: the caller falls back to the full band
            }
            pt.lap("lower diagonals H2D + verify");
        }
        ld.layout = SGV_LAYOUT_DSYM;
        const int rcs = sgv_dsym_ensure_scratch(c, ld);
        pt.lap("kernel scratch");
        return rcs;
    }
    // full band
    const int64_t ldb = round_up(Ml, 32), g0 = c->row_lo;
    float* band = nullptr;
    SGV_CUDA(cudaMalloc(&band, (size_t)(2 * w + 1) * ldb * sizeof(float)));
    ld.band = band;
    ld.owned = true;
    ld.w = w;
    ld.ldb = ldb;
    ld.nnz_stored = (2 * w + 1) * Ml;
    SGV_CUDA(cudaMemsetAsync(band, 0, (size_t)(2 * w + 1) * ldb * sizeof(float), c->stream));
    k_fill_f32<<<592, 256, 0, c->stream>>>(band + w * ldb, Ml, (float)s);
    auto conv = [&](char* st, int* d_offs, int nk) -> int {
        const dim3 grid((unsigned)std::min<int64_t>((Ml + 255) / 256, 2048), (unsigned)nk);
        k_dia_to_band<T><<<grid, blk, 0, c->stream>>>((const T*)st, Ml, d_offs, nk, band, w, ldb, g0, M, s);
        return 0;
    };
    SGV_TRY(stream_diagonals(c, data, sizeof(T), ldd, col0, offsets, all, false, g0, Ml, M, conv));
    ld.layout = SGV_LAYOUT_DIA;
    return 0;
}

extern "C" int sgv_ld_upload_dia(sgv_handle c, int cohort, const void* data, int dtype, int64_t ldd, int64_t col0,
                                 const int64_t* offsets, int ndiag, double s, int layout_hint, int assume_symmetric) {
    SGV_TRY(check_cohort(c, cohort));
    SGV_CHECK(data != nullptr && offsets != nullptr && ndiag > 0, "null / empty DIA arrays");
    SGV_CHECK(dtype == SGV_F32 || dtype == SGV_F64, "bad dtype %d", dtype);
    SGV_CHECK(layout_hint == SGV_LAYOUT_AUTO || layout_hint == SGV_LAYOUT_DIA || layout_hint == SGV_LAYOUT_DSYM,
              "DIA input maps to the band layouts (auto / dia / dsym)");
    SGV_CHECK(c->world == 1 || c->halo, "DIA input needs a halo partition when sharded");
    PhaseTimer pt(c->stream);
    int64_t w = 0;
    for (int k = 0; k < ndiag; ++k) {
        SGV_CHECK(offsets[k] > -c->M && offsets[k] < c->M, "diagonal offset %lld outside the matrix", (long long)offsets[k]);
        w = std::max<int64_t>(w, offsets[k] < 0 ? -offsets[k] : offsets[k]);
        for (int q = 0; q < k; ++q) SGV_CHECK(offsets[q] != offsets[k], "duplicate diagonal offset %lld", (long long)offsets[k]);
    }
    if (c->bandwidth_hint > w) w = c->bandwidth_hint;
    LdMatrix& ld = c->coh[cohort].ld;
    sgv_ld_free(ld);
    pt.lap("free previous LD");
    int layout = layout_hint;
    bool fallback = false;
    if (layout == SGV_LAYOUT_AUTO) {
        layout = sgv_dsym_feasible(w) ? SGV_LAYOUT_DSYM : SGV_LAYOUT_DIA;
        fallback = layout == SGV_LAYOUT_DSYM;
    }
    SGV_CHECK(layout != SGV_LAYOUT_DSYM || sgv_dsym_feasible(w), "DSYM layout infeasible for half-bandwidth %lld", (long long)w);
    SGV_CHECK(layout != SGV_LAYOUT_DIA || sgv_dia_feasible(w), "DIA layout infeasible for half-bandwidth %lld", (long long)w);
    int rc = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        rc = dtype == SGV_F64 ? upload_dia_t<double>(c, ld, (const double*)data, ldd, col0, offsets, ndiag, s, layout, assume_symmetric, w, pt)
                              : upload_dia_t<float>(c, ld, (const float*)data, ldd, col0, offsets, ndiag, s, layout, assume_symmetric, w, pt);
        if (rc != 1) break;
        if (fallback && sgv_dia_feasible(w)) {
            layout = SGV_LAYOUT_DIA;
        } else {
            sgv_set_error("LD matrix is not symmetric: the DSYM layout cannot hold it");
            rc = -1;
            break;
        }
    }
    if (rc != 0) sgv_ld_free(ld);
    pt.lap("DIA upload: rest");
    return rc;
}

extern "C" int sgv_ld_set_bandwidth_hint(sgv_handle c, int64_t w) {
    SGV_CHECK(c != nullptr, "null handle");
    c->bandwidth_hint = w;
    return 0;
}

extern "C" int sgv_ld_info(sgv_handle c, int cohort, int* layout, int64_t* nnz_stored, int64_t* bandwidth,
                           int64_t* nblocks, double* bytes_per_pass_nrhs2) {
    SGV_TRY(check_cohort(c, cohort));
    const LdMatrix& ld = c->coh[cohort].ld;
    if (layout) *layout = ld.layout;
    if (nnz_stored) *nnz_stored = ld.nnz_stored;
    if (bandwidth) *bandwidth = ld.w;
    if (nblocks) *nblocks = ld.nblocks;
    if (bytes_per_pass_nrhs2) {
        // algorithmic bytes of one 2-RHS pass: matrix once + vector pair in + vector pair out
        double b = 32.0 * (double)c->Ml;
        if (ld.layout == SGV_LAYOUT_CSR) b += 8.0 * (double)ld.nnz + 8.0 * (double)(c->Ml + 1);
        else b += 4.0 * (double)ld.nnz_stored;
        *bytes_per_pass_nrhs2 = b;
    }
    return 0;
}
