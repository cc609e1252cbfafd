// Handle lifetime, state vectors, host<->device transfers and the SpMM test / benchmark hooks of
// the C ABI (include/sgvamp_b200.h).
#include <cstdarg>
#include <cstdio>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <thread>
#include "sgv_device.cuh"

static thread_local char g_err[1024] = "";

void sgv_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* sgv_last_error(void) { return g_err; }
extern "C" int sgv_version(void) { return 100; }

__global__ void k_resolve(RedCtx rc);
// Device address of the K x Ml block of r1 vectors (cohort k at k * Ml): the rank-per-cohort mode (src/sgvamp.py:228-233)
// exchanges the cohorts' r1 with ONE collective straight into this block (shard.TorchComm.allgather_r1) instead of K
// host round trips.  The caller orders its collective against the handle's stream (sgv_sync / a shared stream).
extern "C" int sgv_r1_block(sgv_handle c, void** dev_ptr, int64_t* stride) {
    SGV_CHECK(c != nullptr && c->r1_all != nullptr && dev_ptr != nullptr, "handle not configured");
    *dev_ptr = c->r1_all;
    if (stride) *stride = c->Ml;
    return 0;
}

__global__ void k_scale_copy(int64_t M, const double* __restrict__ src, double* __restrict__ dst, double scale);

int sgv_reset_cg_state(sgv_ctx* c) {
    SGV_CUDA(cudaMemsetAsync(c->cg, 0, sizeof(CgState), c->stream));
    SGV_CUDA(cudaMemcpyAsync(&c->cg->band_eps, &c->cg_band_eps, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    return 0;
}

extern "C" int sgv_create(int device, void* stream, sgv_handle* out) {
    SGV_CHECK(out != nullptr, "out is null");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        sgv_set_error("no CUDA device available (%s); libsgvamp_b200 has no CPU fallback",
                      e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return -2;
    }
    SGV_CHECK(device >= 0 && device < ndev, "device %d out of range [0,%d)", device, ndev);
    SGV_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    SGV_CUDA(cudaGetDeviceProperties(&prop, device));
    SGV_CHECK(prop.major >= 10, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
              prop.major, prop.minor);
    sgv_ctx* c = new sgv_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->coop_ok = prop.cooperativeLaunch != 0 && getenv("SGV_NO_COOP") == nullptr;
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        SGV_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    SGV_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    SGV_CUDA(cudaMalloc(&c->counter, 64));
    SGV_CUDA(cudaMemset(c->counter, 0, 64));
    SGV_CUDA(cudaMalloc(&c->cg, sizeof(CgState)));
    SGV_CUDA(cudaMalloc(&c->pubseq, sizeof(unsigned long long)));
    SGV_CUDA(cudaMemset(c->pubseq, 0, sizeof(unsigned long long)));
    if (const char* e = getenv("SGV_CG_BAND")) c->cg_band_eps = atof(e);
    SGV_TRY(sgv_reset_cg_state(c));
    SGV_CUDA(cudaStreamSynchronize(c->stream));
    SGV_CUDA(cudaMallocHost(&c->cg_host, sizeof(CgState)));
    memset(c->cg_host, 0, sizeof(CgState));
    SGV_CUDA(cudaMallocHost(&c->host_scal, 64 * sizeof(double)));
    SGV_CUDA(cudaEventCreate(&c->ev_a));
    SGV_CUDA(cudaEventCreate(&c->ev_b));
    SGV_CUDA(cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming));
    {   // load every kernel now (lazy loading inside the solver loop synchronises the device)
        cudaFuncAttributes fa;
        SGV_CUDA(cudaFuncGetAttributes(&fa, k_resolve));
        SGV_CUDA(cudaFuncGetAttributes(&fa, k_scale_copy));
        SGV_TRY(sgv_preload_spmm());
        SGV_TRY(sgv_preload_dsym());
        SGV_TRY(sgv_preload_psym());
        SGV_TRY(sgv_preload_vamp());
    }
    *out = c;
    return 0;
}

static void detach_peers(sgv_ctx* c) {
    for (int q = 0; q < SGV_MAX_RANKS; ++q) {
        if (c->peer[q].ipc && c->peer[q].base) cudaIpcCloseMemHandle(c->peer[q].base);
        c->peer[q] = PeerView();
    }
}

static void free_vectors(sgv_ctx* c) {
    for (int k = 0; k < SGV_MAX_K; ++k) {
        Cohort& co = c->coh[k];
        sgv_ld_free(co.ld);
        cudaFree(co.xty);
        cudaFree(co.r2);
        cudaFree(co.xhat2);
        cudaFree(co.sig);
        cudaFree(co.probe);
        cudaFree(co.rxs);
        co = Cohort();
    }
    detach_peers(c);
    cudaFree(c->arena);
    cudaFree(c->bb);
    cudaFree(c->vfull);
    cudaFree(c->probe_b);
    cudaFree(c->em_cache);
    c->em_cache = nullptr;
    c->em_cache_cap = 0;
    c->vfull = nullptr;
    c->probe_b = nullptr;
    c->arena = nullptr;
    c->bb = c->qq = c->xx = c->rr = c->pp[0] = c->pp[1] = c->rr2[0] = c->rr2[1] = c->qq2[0] = c->qq2[1] = nullptr;
    cudaFree(c->r1_all);
    cudaFree(c->xhat1);
    cudaFree(c->truth);
    c->r1_all = c->xhat1 = c->truth = nullptr;
    for (int i = 0; i < sgv_ctx::NSNAP; ++i) {
        if (c->snap[i]) {
            cudaFree(c->snap[i]);
            cudaEventDestroy(c->snap_ev[i]);
            c->snap[i] = nullptr;
            c->snap_ev[i] = nullptr;
        }
    }
    c->snap_next = 0;
}

extern "C" int sgv_destroy(sgv_handle c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->copy_stream);
    if (c->dsp_dbg) {   // SGV_DS_DEBUG: where a CG step of the whole-solve kernel spends its time (range 0 / last CTA), per rank
        unsigned long long d[16];
        if (cudaMemcpy(d, c->dsp_dbg, sizeof(d), cudaMemcpyDeviceToHost) == cudaSuccess && d[7] > 0) {
            const double n = (double)d[7];
            fprintf(stderr, "[sgv solve clock] rank %d/%d steps %llu | us per step: stage %.1f tiles %.1f headfix %.1f wait(range0) %.1f | "
                            "last CTA: reduce %.1f exchange %.1f\n", c->rank, c->world, d[7], d[0] / n / 1e3, d[1] / n / 1e3,
                    d[2] / n / 1e3, d[3] / n / 1e3, d[4] / n / 1e3, d[5] / n / 1e3);
        }
        cudaFree(c->dsp_dbg);
    }
    free_vectors(c);
    cudaFree(c->ypart);
    cudaFree(c->ypartT);
    cudaFree(c->ds_ypart);
    cudaFree(c->ds_tails);
    cudaFree(c->dsp_yhead);
    cudaFree(c->dsp_tails);
    cudaFree(c->dsp_flags);
    cudaFree(c->partials);
    cudaFree(c->counter);
    cudaFree(c->cg);
    cudaFree(c->pubseq);
    cudaFree(c->stage);
    cudaFreeHost(c->cg_host);
    cudaFreeHost(c->host_scal);
    cudaEventDestroy(c->ev_a);
    cudaEventDestroy(c->ev_b);
    cudaEventDestroy(c->ev_copy);
    for (cudaEvent_t e : c->prof_ev) cudaEventDestroy(e);
    for (int i = 0; i < sgv_ctx::NLOG; ++i) {
        if (c->log_host[i]) cudaFreeHost(c->log_host[i]);
        if (c->log_ev[i]) cudaEventDestroy(c->log_ev[i]);
        if (c->probe_pin[i]) cudaFreeHost(c->probe_pin[i]);
    }
    cudaStreamDestroy(c->copy_stream);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

extern "C" int sgv_sync(sgv_handle c) {
    SGV_CHECK(c != nullptr, "null handle");
    SGV_CUDA(cudaStreamSynchronize(c->stream));
    SGV_CUDA(cudaStreamSynchronize(c->copy_stream));
    return 0;
}

extern "C" int64_t sgv_launch_count(sgv_handle c) { return c ? c->launches : -1; }

extern "C" int sgv_profile(sgv_handle c, int enable) {
    SGV_CHECK(c != nullptr, "null handle");
    SGV_CUDA(cudaStreamSynchronize(c->stream));
    c->prof = enable != 0;
    if (enable) c->prof_n = 0;
    return 0;
}

extern "C" int sgv_profile_read(sgv_handle c, double* total_ms, int64_t* launches) {
    SGV_CHECK(c != nullptr && total_ms && launches, "null argument");
    SGV_CUDA(cudaStreamSynchronize(c->stream));
    double tot = 0.0;
    for (size_t i = 0; i + 1 < c->prof_n; i += 2) {
        float ms = 0.f;
        SGV_CUDA(cudaEventElapsedTime(&ms, c->prof_ev[i], c->prof_ev[i + 1]));
        tot += ms;
    }
    *total_ms = tot;
    *launches = (int64_t)(c->prof_n / 2);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// cross-rank reduction plumbing
// ---------------------------------------------------------------------------------------------
// One warp; lane q waits for rank q's partial sums of reduction `seq`, lane 0 combines the rows in
// rank order (bit-identical on all ranks) and applies the state transition.  A bounded spin: on
// time-out the error flag is raised and both CG columns are marked done so that nothing hangs.
__global__ void k_resolve(RedCtx rc) {
    // nobody published: the producers exited early too
    if (rc.skip_if_done == SKIP_CG_DONE && rc.st->done[0] && rc.st->done[1]) return;
    if (rc.skip_if_done == SKIP_EM_DONE && rc.st->em.done) return;
    const unsigned long long seq = red_next_seq(rc);
    __syncwarp();
    resolve_warp(rc, threadIdx.x, seq);
}

RedCtx sgv_red_begin(sgv_ctx* c, int kind, int nv, int off, int maxit, int x0_zero, int is_min) {
    RedCtx rc;
    memset(&rc, 0, sizeof(rc));
    rc.partials = c->partials;
    rc.counter = c->counter;
    rc.st = c->cg;
    rc.world = c->world;
    rc.rank = c->rank;
    rc.seq = ++c->seq;
    rc.pubseq = c->pubseq;
    rc.inline_resolve = c->world > 1 && !c->host_barrier;
    for (int q = 0; q < c->world; ++q) rc.inbox[q] = reinterpret_cast<Inbox*>(c->peer[q].base);
    rc.ap.kind = kind;
    rc.ap.nv = nv;
    rc.ap.off = off;
    rc.ap.maxit = maxit;
    rc.ap.x0_zero = x0_zero;
    rc.ap.is_min = is_min;
    return rc;
}

int sgv_red_end(sgv_ctx* c, const RedCtx& rc) {
    if (c->world > 1 && c->host_barrier) {
        // shared-GPU mode: wait on the host until every rank's reducing kernel has finished
        SGV_CUDA(cudaStreamSynchronize(c->stream));
        const unsigned long long my = ++c->host_seq;
        const auto t0 = std::chrono::steady_clock::now();
        for (int q = 0; q < c->world; ++q) {
            if (q == c->rank) continue;
            SGV_CHECK(c->peer_ctx[q] != nullptr, "host-barrier mode needs locally attached peers");
            while (c->peer_ctx[q]->host_seq.load(std::memory_order_acquire) < my) {
                std::this_thread::sleep_for(std::chrono::microseconds(20));
                SGV_CHECK(std::chrono::steady_clock::now() - t0 < std::chrono::seconds(60),
                          "host barrier timed out waiting for rank %d (rank %d of %d)", q, c->rank, c->world);
            }
        }
    }
    if (c->world > 1 && !rc.inline_resolve) {
        k_resolve<<<1, 32, 0, c->stream>>>(rc);
        c->launches++;
        SGV_CUDA(cudaGetLastError());
    }
    return 0;
}

extern "C" int sgv_configure_part(sgv_handle c, int64_t M, int K, int rank, int world, int64_t row_lo, int64_t row_hi,
                                  int halo) {
    SGV_CHECK(c != nullptr, "null handle");
    SGV_CHECK(M > 0, "M must be positive");
    SGV_CHECK(K >= 1 && K <= SGV_MAX_K, "K=%d outside [1,%d]", K, SGV_MAX_K);
    SGV_CHECK(world >= 1 && world <= SGV_MAX_RANKS && rank >= 0 && rank < world, "bad rank/world %d/%d", rank, world);
    SGV_CHECK(row_lo >= 0 && row_hi > row_lo && row_hi <= M, "bad row range [%lld,%lld)", (long long)row_lo, (long long)row_hi);
    SGV_CUDA(cudaSetDevice(c->device));
    SGV_CUDA(cudaStreamSynchronize(c->stream));
    SGV_CUDA(cudaStreamSynchronize(c->copy_stream));
    free_vectors(c);
    c->M = M;
    c->Ml = row_hi - row_lo;
    c->row_lo = row_lo;
    c->rank = rank;
    c->world = world;
    SGV_CHECK(halo >= 0 && halo <= 2, "halo must be 0 (block-diagonal shards), 1 (banded halos) or 2 (dense rows)");
    c->halo = halo == 1;
    c->rowpart = halo == 2;
    c->K = K;
    c->seq = 0;
    c->vamp_begun = false;
    if (c->probe_pin_bytes < (int64_t)K * (row_hi - row_lo)) c->probe_pin_bytes = 0;   // re-allocated by sgv_vamp_begin
    c->host_seq.store(0);
    c->host_barrier = false;
    for (int q = 0; q < SGV_MAX_RANKS; ++q) c->peer_ctx[q] = nullptr;
    c->prior.K = K;
    for (int k = 0; k < K; ++k) c->prior.a[k] = 1.0 / K;
    const int64_t Ml = c->Ml;
    const size_t vb = (size_t)Ml * sizeof(double), v2 = (size_t)Ml * sizeof(double2);
    SGV_CUDA(cudaMalloc(&c->r1_all, vb * K));
    SGV_CUDA(cudaMalloc(&c->xhat1, vb));
    SGV_CUDA(cudaMemsetAsync(c->r1_all, 0, vb * K, c->stream));
    SGV_CUDA(cudaMemsetAsync(c->xhat1, 0, vb, c->stream));
    for (int k = 0; k < K; ++k) {
        Cohort& co = c->coh[k];
        co.r1 = c->r1_all + (size_t)k * Ml;
        SGV_CUDA(cudaMalloc(&co.xty, vb));
        SGV_CUDA(cudaMalloc(&co.r2, vb));
        SGV_CUDA(cudaMalloc(&co.xhat2, vb));
        SGV_CUDA(cudaMalloc(&co.sig, vb));
        SGV_CUDA(cudaMalloc(&co.probe, Ml));
        SGV_CUDA(cudaMalloc(&co.rxs, v2));
        SGV_CUDA(cudaMemsetAsync(co.rxs, 0, v2, c->stream));
        double* z[] = {co.xty, co.r2, co.xhat2, co.sig};
        for (double* p : z) SGV_CUDA(cudaMemsetAsync(p, 0, vb, c->stream));
    }
    // symmetric arena: [Inbox | xx | rr | pp0 | pp1]
    c->arena_bytes = arena_size(Ml);
    SGV_CUDA(cudaMalloc(&c->arena, c->arena_bytes));
    SGV_CUDA(cudaMemsetAsync(c->arena, 0, c->arena_bytes, c->stream));
    SGV_CUDA(cudaMemsetAsync(c->pubseq, 0, sizeof(unsigned long long), c->stream));   // slot numbering restarts with the zeroed inbox
    c->xx = reinterpret_cast<double2*>(c->arena + arena_off_xx(Ml));
    for (int i = 0; i < 2; ++i) {
        c->rr2[i] = reinterpret_cast<double2*>(c->arena + arena_off_rr(Ml, i));
        c->pp[i] = reinterpret_cast<double2*>(c->arena + arena_off_pp(Ml, i));
        c->qq2[i] = reinterpret_cast<double2*>(c->arena + arena_off_qq(Ml, i));
    }
    c->rr = c->rr2[1];   // where the set-up kernel puts r_0 (CG step 0 reads buffer 1, writes buffer 0)
    c->qq = c->qq2[0];
    c->peer[rank].base = c->arena;
    c->peer[rank].Ml = Ml;
    c->peer[rank].ipc = false;
    SGV_CUDA(cudaMalloc(&c->bb, v2));
    SGV_CUDA(cudaMemsetAsync(c->bb, 0, v2, c->stream));
    SGV_CUDA(cudaMalloc(&c->probe_b, Ml));
    if (c->rowpart) {
        SGV_CUDA(cudaMalloc(&c->vfull, (size_t)M * sizeof(double2)));
        SGV_CUDA(cudaMemsetAsync(c->vfull, 0, (size_t)M * sizeof(double2), c->stream));
    }
    SGV_TRY(sgv_reset_cg_state(c));
    for (int i = 0; i < sgv_ctx::NSNAP; ++i) {   // allocated up front: cudaMalloc inside the loop would synchronise
        SGV_CUDA(cudaMalloc(&c->snap[i], vb));
        SGV_CUDA(cudaEventCreateWithFlags(&c->snap_ev[i], cudaEventDisableTiming));
    }
    // Everything the solver loop needs is allocated here: a cudaMalloc / cudaFree inside the loop is a
    // device-wide synchronisation, which would deadlock against another rank's resolve kernel spinning
    // on the same device (ranks sharing a GPU) and stall the pipeline otherwise.
    SGV_CUDA(cudaMalloc(&c->truth, vb));
    SGV_CUDA(cudaMemsetAsync(c->truth, 0, vb, c->stream));
    c->truth_set = false;
    SGV_TRY(sgv_ensure_partials(c, std::max<int64_t>((int64_t)c->sm_count * 16, (Ml + 7) / 8 + 1)));
    SGV_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int sgv_configure(sgv_handle c, int64_t M, int K) { return sgv_configure_part(c, M, K, 0, 1, 0, M, 0); }

// ---- peer attachment: every rank maps every other rank's arena ----
extern "C" int sgv_ipc_export(sgv_handle c, void* handle64) {
    SGV_CHECK(c != nullptr && c->arena != nullptr && handle64 != nullptr, "handle not configured");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "unexpected IPC handle size");
    SGV_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), c->arena));
    return 0;
}

extern "C" int sgv_ipc_import(sgv_handle c, int peer_rank, const void* handle64, int64_t peer_Ml) {
    SGV_CHECK(c != nullptr && c->arena != nullptr, "handle not configured");
    SGV_CHECK(peer_rank >= 0 && peer_rank < c->world && peer_rank != c->rank, "bad peer rank %d", peer_rank);
    SGV_CUDA(cudaSetDevice(c->device));
    cudaIpcMemHandle_t hdl;
    memcpy(&hdl, handle64, sizeof(hdl));
    void* p = nullptr;
    SGV_CUDA(cudaIpcOpenMemHandle(&p, hdl, cudaIpcMemLazyEnablePeerAccess));
    c->peer[peer_rank].base = static_cast<char*>(p);
    c->peer[peer_rank].Ml = peer_Ml;
    c->peer[peer_rank].ipc = true;
    return 0;
}

// same-process variant (one host thread per GPU): direct pointers + cudaDeviceEnablePeerAccess
extern "C" int sgv_peer_attach_local(sgv_handle c, int peer_rank, sgv_handle other) {
    SGV_CHECK(c != nullptr && other != nullptr && c->arena && other->arena, "handles not configured");
    SGV_CHECK(peer_rank >= 0 && peer_rank < c->world && peer_rank != c->rank, "bad peer rank %d", peer_rank);
    SGV_CUDA(cudaSetDevice(c->device));
    if (other->device != c->device) {
        int can = 0;
        SGV_CUDA(cudaDeviceCanAccessPeer(&can, c->device, other->device));
        SGV_CHECK(can, "device %d cannot access device %d", c->device, other->device);
        cudaError_t e = cudaDeviceEnablePeerAccess(other->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) SGV_CUDA(e);
        cudaGetLastError();
    }
    c->peer[peer_rank].base = other->arena;
    c->peer[peer_rank].Ml = other->Ml;
    c->peer[peer_rank].ipc = false;
    c->peer_ctx[peer_rank] = other;
    return 0;
}

extern "C" int sgv_set_host_barrier(sgv_handle c, int enable) {
    SGV_CHECK(c != nullptr, "null handle");
    c->host_barrier = enable != 0;
    return 0;
}

extern "C" int sgv_device_id(sgv_handle c, char* pci_bus_id, int len) {
    SGV_CHECK(c != nullptr && pci_bus_id != nullptr && len >= 16, "bad arguments");
    SGV_CUDA(cudaDeviceGetPCIBusId(pci_bus_id, len, c->device));
    return 0;
}

extern "C" int sgv_partition_info(sgv_handle c, int64_t* M, int64_t* Ml, int64_t* row_lo, int* rank, int* world) {
    SGV_CHECK(c != nullptr, "null handle");
    if (M) *M = c->M;
    if (Ml) *Ml = c->Ml;
    if (row_lo) *row_lo = c->row_lo;
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    return 0;
}

static int vec_ptr(sgv_ctx* c, int cohort, int which, double** p) {
    SGV_CHECK(c != nullptr && c->M > 0, "handle not configured");
    SGV_CHECK(cohort >= 0 && cohort < c->K, "cohort index %d out of range [0,%d)", cohort, c->K);
    SGV_CUDA(cudaSetDevice(c->device));
    Cohort& co = c->coh[cohort];
    switch (which) {
        case SGV_VEC_XHAT1: *p = c->xhat1; break;
        case SGV_VEC_R1: *p = co.r1; break;
        case SGV_VEC_XHAT2: *p = co.xhat2; break;
        case SGV_VEC_SIGMA2U: *p = co.sig; break;
        case SGV_VEC_R2: *p = co.r2; break;
        case SGV_VEC_XTY: *p = co.xty; break;
        default: sgv_set_error("unknown vector id %d", which); return -1;
    }
    return 0;
}

extern "C" int sgv_set_xty(sgv_handle c, int cohort, const double* r) {
    double* p;
    SGV_TRY(vec_ptr(c, cohort, SGV_VEC_XTY, &p));
    SGV_CHECK(r != nullptr, "r is null");
    SGV_CUDA(cudaMemcpyAsync(p, r, c->Ml * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    SGV_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int sgv_reset_state(sgv_handle c) {
    SGV_CHECK(c != nullptr && c->M > 0, "handle not configured");
    SGV_CUDA(cudaSetDevice(c->device));
    const size_t vb = (size_t)c->Ml * sizeof(double);
    SGV_CUDA(cudaMemsetAsync(c->xhat1, 0, vb, c->stream));
    for (int k = 0; k < c->K; ++k) {
        Cohort& co = c->coh[k];
        SGV_CUDA(cudaMemcpyAsync(co.r1, co.xty, vb, cudaMemcpyDeviceToDevice, c->stream));   // r1 <- r  (:204)
        SGV_CUDA(cudaMemsetAsync(co.xhat2, 0, vb, c->stream));
        SGV_CUDA(cudaMemsetAsync(co.sig, 0, vb, c->stream));
        SGV_CUDA(cudaMemsetAsync(co.r2, 0, vb, c->stream));
        SGV_CUDA(cudaMemsetAsync(co.rxs, 0, (size_t)c->Ml * sizeof(double2), c->stream));
        co.rxs_valid = true;
    }
    return 0;
}

extern "C" int sgv_get_vec(sgv_handle c, int cohort, int which, double* dst) {
    double* p;
    SGV_TRY(vec_ptr(c, cohort, which, &p));
    SGV_CUDA(cudaMemcpyAsync(dst, p, c->Ml * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SGV_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int sgv_set_vec(sgv_handle c, int cohort, int which, const double* src) {
    double* p;
    SGV_TRY(vec_ptr(c, cohort, which, &p));
    SGV_CUDA(cudaMemcpyAsync(p, src, c->Ml * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    SGV_CUDA(cudaStreamSynchronize(c->stream));
    if (which == SGV_VEC_XHAT2 || which == SGV_VEC_SIGMA2U) c->coh[cohort].rxs_valid = false;
    return 0;
}

__global__ void k_scale_copy(int64_t M, const double* __restrict__ src, double* __restrict__ dst, double scale) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (int64_t)gridDim.x * blockDim.x)
        dst[j] = src[j] * scale;
}

// Snapshot (scaled) into a device staging slot on the compute stream, then copy to pinned host
// memory on the copy stream so that the transfer overlaps the following kernels.
extern "C" int sgv_get_vec_async(sgv_handle c, int cohort, int which, double scale, double* pinned_dst) {
    double* p;
    SGV_TRY(vec_ptr(c, cohort, which, &p));
    const int slot = c->snap_next;
    c->snap_next = (c->snap_next + 1) % sgv_ctx::NSNAP;
    SGV_CUDA(cudaStreamWaitEvent(c->stream, c->snap_ev[slot], 0));   // previous read-back of this slot is done
    double* snap = c->snap[slot];
    const unsigned grid = (unsigned)std::min<int64_t>((c->Ml + 255) / 256, (int64_t)c->sm_count * 8);
    k_scale_copy<<<grid, 256, 0, c->stream>>>(c->Ml, p, snap, scale);
    c->launches++;
    SGV_CUDA(cudaEventRecord(c->ev_copy, c->stream));
    SGV_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_copy, 0));
    SGV_CUDA(cudaMemcpyAsync(pinned_dst, snap, c->Ml * sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream));
    SGV_CUDA(cudaEventRecord(c->snap_ev[slot], c->copy_stream));
    return 0;
}

extern "C" int sgv_wait_copies(sgv_handle c) {
    SGV_CHECK(c != nullptr, "null handle");
    SGV_CUDA(cudaStreamSynchronize(c->copy_stream));
    return 0;
}

extern "C" int sgv_pinned_alloc(sgv_handle c, int64_t bytes, void** out) {
    SGV_CHECK(c != nullptr && out != nullptr, "null argument");
    SGV_CUDA(cudaSetDevice(c->device));   // callable from any host thread: never touch (and initialise) another device
    SGV_CUDA(cudaMallocHost(out, bytes));
    return 0;
}

extern "C" int sgv_pinned_free(sgv_handle c, void* p) {
    SGV_CHECK(c != nullptr, "null handle");
    SGV_CUDA(cudaFreeHost(p));
    return 0;
}

extern "C" int sgv_set_prior(sgv_handle c, int L, double lam, const double* omegas, const double* sigmas) {
    SGV_CHECK(c != nullptr, "null handle");
    SGV_CHECK(L >= 2 && L <= SGV_MAX_L, "L=%d outside [2,%d]", L, SGV_MAX_L);
    c->prior.L = L;
    c->prior.lam = lam;
    for (int l = 0; l < L - 1; ++l) {
        c->prior.omegas[l] = omegas[l];
        c->prior.sigmas[l] = sigmas[l];
    }
    return 0;
}

extern "C" int sgv_set_weights(sgv_handle c, const double* a) {
    SGV_CHECK(c != nullptr && c->K > 0, "handle not configured");
    for (int k = 0; k < c->K; ++k) c->prior.a[k] = a[k];
    return 0;
}

// ---------------------------------------------------------------------------------------------
// SpMM hooks
// ---------------------------------------------------------------------------------------------
extern "C" int sgv_spmm(sgv_handle c, int cohort, const double* X, double* Y, int nrhs, double alpha, double beta) {
    SGV_CHECK(c != nullptr && c->M > 0, "handle not configured");
    SGV_CUDA(cudaSetDevice(c->device));
    SGV_CHECK(cohort >= 0 && cohort < c->K, "cohort out of range");
    SGV_CHECK(nrhs == 1 || nrhs == 2, "nrhs must be 1 or 2");
    Cohort& co = c->coh[cohort];
    const int64_t M = c->Ml;
    std::vector<double2> h(M);
    for (int64_t i = 0; i < M; ++i) h[i] = make_double2(X[i], nrhs == 2 ? X[M + i] : 0.0);
    SGV_CUDA(cudaMemcpyAsync(c->pp[0], h.data(), M * sizeof(double2), cudaMemcpyHostToDevice, c->stream));
    SGV_TRY(sgv_launch_spmm(c, co, EPI_PLAIN, VEC_PP0, c->qq, alpha, beta, 0, 0));
    SGV_CUDA(cudaMemcpyAsync(h.data(), c->qq, M * sizeof(double2), cudaMemcpyDeviceToHost, c->stream));
    SGV_CUDA(cudaStreamSynchronize(c->stream));
    for (int64_t i = 0; i < M; ++i) {
        Y[i] = h[i].x;
        if (nrhs == 2) Y[M + i] = h[i].y;
    }
    return 0;
}

// In multi-rank runs the vector upload of sgv_spmm must be complete on every rank before any rank's
// kernel reads its neighbours' halos: sgv_spmm_stage uploads, the caller synchronises the ranks
// (barrier), sgv_spmm_run multiplies.
extern "C" int sgv_spmm_stage(sgv_handle c, const double* X, int nrhs) {
    SGV_CHECK(c != nullptr && c->M > 0, "handle not configured");
    SGV_CUDA(cudaSetDevice(c->device));
    const int64_t M = c->Ml;
    std::vector<double2> h(M);
    for (int64_t i = 0; i < M; ++i) h[i] = make_double2(X[i], nrhs == 2 ? X[M + i] : 0.0);
    SGV_CUDA(cudaMemcpyAsync(c->pp[0], h.data(), M * sizeof(double2), cudaMemcpyHostToDevice, c->stream));
    SGV_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int sgv_spmm_run(sgv_handle c, int cohort, double* Y, int nrhs, double alpha, double beta) {
    SGV_CHECK(c != nullptr && c->M > 0, "handle not configured");
    SGV_CUDA(cudaSetDevice(c->device));
    SGV_CHECK(cohort >= 0 && cohort < c->K, "cohort out of range");
    Cohort& co = c->coh[cohort];
    const int64_t M = c->Ml;
    std::vector<double2> h(M);
    SGV_TRY(sgv_launch_spmm(c, co, EPI_PLAIN, VEC_PP0, c->qq, alpha, beta, 0, 0));
    SGV_CUDA(cudaMemcpyAsync(h.data(), c->qq, M * sizeof(double2), cudaMemcpyDeviceToHost, c->stream));
    SGV_CUDA(cudaStreamSynchronize(c->stream));
    for (int64_t i = 0; i < M; ++i) {
        Y[i] = h[i].x;
        if (nrhs == 2) Y[M + i] = h[i].y;
    }
    return 0;
}

extern "C" int sgv_spmm_bench(sgv_handle c, int cohort, int reps, float* ms_per_launch) {
    SGV_CHECK(c != nullptr && c->M > 0, "handle not configured");
    SGV_CHECK(cohort >= 0 && cohort < c->K, "cohort out of range");
    SGV_CHECK(reps > 0 && ms_per_launch, "bad arguments");
    SGV_CHECK(c->world == 1, "sgv_spmm_bench is a single-rank hook");
    Cohort& co = c->coh[cohort];
    // the CG-shaped pass: q = gamw*(R p) + gam2*p with the p.q dots, on whatever the vectors hold
    for (int i = 0; i < 3; ++i) SGV_TRY(sgv_launch_spmm(c, co, EPI_Q, VEC_PP0, c->qq, 1.5, 0.25, 0, 0));
    SGV_CUDA(cudaEventRecord(c->ev_a, c->stream));
    for (int i = 0; i < reps; ++i) SGV_TRY(sgv_launch_spmm(c, co, EPI_Q, VEC_PP0, c->qq, 1.5, 0.25, 0, 0));
    SGV_CUDA(cudaEventRecord(c->ev_b, c->stream));
    SGV_CUDA(cudaEventSynchronize(c->ev_b));
    float ms = 0.f;
    SGV_CUDA(cudaEventElapsedTime(&ms, c->ev_a, c->ev_b));
    *ms_per_launch = ms / reps;
    return 0;
}
