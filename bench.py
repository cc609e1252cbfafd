#!/usr/bin/env python
"""Benchmark of the sgVAMP hot path: VAMP iterations / second at M markers with banded LD
(BASELINE.json metric; configs[4]: M=1M banded LD row-partitioned over 1/2/4/8 B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--M M] [--w W]

A "step" is one VAMP iteration (denoiser -> prior EM -> 2-RHS CG LMMSE -> Hutchinson -> gamw)
on synthetic genotype-derived banded LD.  The run performs W warm-up iterations followed by K
timed iterations of one continuing VAMP trajectory; the timed region is bracketed by CUDA events
on the solver's stream (max over ranks).  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(REPO, "sgvamp-py_b200")
for p in (REPO, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

H2, LAM_TRUE, S_REG, N_LD = 0.5, 0.01, 0.1, 4096


def n_gwas(M):
    """GWAS sample size of the synthetic cohort.  The gamw update (src/sgvamp.py:352-363) assumes
    r = X^T y for an N x M design, so N must be at least rank(R) = M for a full-rank banded R."""
    return 2 * int(M)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--M", type=int, default=1_000_000)
    ap.add_argument("--w", type=int, default=500)
    ap.add_argument("--cpu-sample-M", type=int, default=80_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-format", default="dia", choices=["dia", "csr"],
                    help="host container of the LD matrix in the end-to-end leg (scipy DIA arrays or CSR)")
    ap.add_argument("--seed", type=int, default=5)
    ap.add_argument("--config", default="c5", choices=["c1", "c2", "c3", "c4", "c5"],
                    help="BASELINE.json configuration: c5 (default, the headline metric) = M=1M banded LD row-partitioned "
                         "over the GPUs; c1..c4 = the dense / block-diagonal shapes on one GPU (bench_configs.py)")
    return ap.parse_args()


def vamp_params(M):
    cm = max(1, int(M * LAM_TRUE))
    return dict(prior_vars=[0.0, H2 / cm], prior_probs=[0.99, 0.01], rho=0.2, gamw=2.0, gam1=1e-6,
                cg_maxit=500, em_prior_maxit=100, learn_gamw=True, lmmse_damp=False, prior_update="em",
                update_prior_from=1)


def make_probes(iterations, M, seed):
    rng = np.random.RandomState(1234 + seed)
    return (rng.binomial(p=0.5, n=1, size=(1, iterations, M)) * 2 - 1).astype(np.int8)


# ------------------------------------------------------------------------------------------------
# synthetic banded LD + XTy on the device (torch as a data-generation utility)
# ------------------------------------------------------------------------------------------------
def build_problem(torch, M, w, seed, dev, lo=0, hi=None, ext=0, keep_full=True, use_library=True):
    """Rows [lo, hi) of the workload.  Returns the symmetric half band of Rused in the DSYM layout
    (fp32, device; with `ext` leading extension rows for ranks > 0, sgv_ld_adopt_dsym), the full band
    of the rows [lo-ext, hi) (for the host-side end-to-end leg; None unless keep_full; the extension rows hold only their
    couplings to the own rows), r (host), x0 (host, global)."""
    import ldgen
    t0 = time.time()
    hi = M if hi is None else hi
    Ml = hi - lo
    glo = lo - ext                                  # first generated row (extension rows included)
    assert glo >= 0
    n = hi - glo
    # LD by the library's own construction kernel (sgv_ld_build_banded: exact integer Gram sums of the int8 genotypes,
    # standardisation + Bartlett taper + Rused = (1-s) R + s I in its epilogue, tiled half band as output); torch only
    # synthesises the genotypes and the noise
    Dp = (w + 1 + 3) // 4 * 4
    if use_library:
        import sgv_native as nat
        part = lo > 0 or hi < M
        U, ldb, ext_lib, noise_own = ldgen.banded_dsym_library(torch, nat, M, w, lo, hi, 1 if ext > 0 else 0, 2 if part else 1,
                                                                seed, dev, S_REG, N_ld=N_LD)
        assert ext_lib == ext, (ext_lib, ext)
        Ud = ldgen.dsym_untile(torch, U, Dp, ldb)[: w + 1, :n]
        band = torch.zeros((2 * w + 1, n), device=dev, dtype=torch.float32)   # full band of the rows [glo, hi) (r, host legs)
        band[w:, :] = Ud
        band[w] *= 2.0                              # the layout stores half of the diagonal
        for d in range(1, w + 1):
            band[w - d, d:] = Ud[d, : n - d]
        del Ud
    else:
        # the same workload from torch matmuls (reference arm: nothing of this library on its path; CPU-only boxes)
        band, noise = ldgen.banded_dia_device(torch, M, w, glo, hi, seed, dev, N_ld=N_LD)   # (2w+1) x n
        for d in range(1, w + 1):                   # exactly symmetric: the lower triangle mirrors the upper one
            band[w - d, d:] = band[w + d, : n - d]
        band *= (1.0 - S_REG)                       # Rused = (1-s) R + s I  (src/main.py:265)
        band[w, :] += S_REG
        noise_own = noise[ext:]
        ldb = (n + 127) // 128 * 128
        U = None
    x0_host = ldgen.causal_effects(M, n_gwas(M), LAM_TRUE, H2, seed)
    # r = Rused x0 + n,  n ~ N(0, (1-h2) Rused): the summary-statistic form of the reference recipe
    # (simulation/sim_gen_phen_mult.py:39-55: r = X^T y, R = X^T X  =>  r ~ N(R x0, (1-h2) R))
    xp_host = np.zeros(Ml + 2 * w)
    a0, a1 = max(0, lo - w), min(M, hi + w)
    xp_host[a0 - (lo - w): a1 - (lo - w)] = x0_host[a0:a1]
    xp = torch.from_numpy(xp_host).to(dev)
    r = torch.zeros(Ml, device=dev, dtype=torch.float64)
    for d in range(2 * w + 1):
        r += band[d, ext:].to(torch.float64) * xp[d:d + Ml]
    g = torch.Generator(device=dev)
    g.manual_seed(seed * 31 + 17)
    z = torch.randn((M,), generator=g, device=dev, dtype=torch.float64)[lo:hi]
    r += float(np.sqrt(1.0 - H2)) * (float(np.sqrt(1.0 - S_REG)) * noise_own + float(np.sqrt(S_REG)) * z)
    full = band if keep_full else None              # full band of the rows [glo, hi) (host-side legs)
    del band
    if dev.type == "cuda":
        torch.cuda.synchronize()
    return U, ldb, full, r.cpu().numpy(), x0_host, time.time() - t0


def band_to_host_csr(torch, band, M, w, lo=0, hi=None, pinned=True):
    """CSR (fp32 data, int32 GLOBAL column indices) of rows [lo, hi) of the banded matrix in (pinned)
    host memory - the host-side input of the end-to-end leg."""
    hi = M if hi is None else hi
    Ml = hi - lo
    rows_all = np.arange(lo, hi, dtype=np.int64)
    cnt = np.minimum(rows_all + w, M - 1) - np.maximum(rows_all - w, 0) + 1
    nnz = int(cnt.sum())
    data = torch.empty(nnz, dtype=torch.float32, pin_memory=pinned)
    idx = torch.empty(nnz, dtype=torch.int32, pin_memory=pinned)
    indptr = np.zeros(Ml + 1, dtype=np.int64)
    indptr[1:] = np.cumsum(cnt)
    dd = torch.arange(2 * w + 1, device=band.device)
    pos = 0
    step = 32768
    for i0 in range(0, Ml, step):
        i1 = min(Ml, i0 + step)
        rows = torch.arange(lo + i0, lo + i1, device=band.device)
        cols = rows[:, None] + (dd[None, :] - w)
        ok = (cols >= 0) & (cols < M)
        vals = band[:, i0:i1].t()[ok]
        cc = cols[ok].to(torch.int32)
        n = vals.numel()
        data[pos:pos + n].copy_(vals)
        idx[pos:pos + n].copy_(cc)
        pos += n
    assert pos == nnz
    torch.cuda.synchronize()
    import scipy.sparse
    R = scipy.sparse.csr_matrix((data.numpy(), idx.numpy(), indptr.astype(np.int32) if nnz < 2**31 else indptr),
                                shape=(Ml, M))
    R.has_canonical_format = True
    return R, (data, idx)


def band_to_host_dia(torch, band, M, w, glo, hi, pinned=True):
    """The rows [glo, hi) of the banded matrix as a column window of scipy's DIA format in (pinned) host
    memory: data[k, j - col0] = R[j - off_k, j], off_k = k - w  (band[k, t] = R[glo + t, glo + t + off_k])."""
    n = hi - glo
    col0 = max(0, glo - w)
    ldd = min(M, hi + w) - col0
    data = torch.zeros((2 * w + 1, ldd), device=band.device, dtype=torch.float32)
    for k in range(2 * w + 1):
        sh = glo + (k - w) - col0                       # column of row glo inside the window
        t0, t1 = max(0, -sh), min(n, ldd - sh)
        if t1 > t0:
            data[k, t0 + sh: t1 + sh] = band[k, t0:t1]
    host = torch.empty((2 * w + 1, ldd), dtype=torch.float32, pin_memory=pinned)
    host.copy_(data)
    if band.is_cuda:
        torch.cuda.synchronize()
    return host, np.arange(-w, w + 1, dtype=np.int64), col0


class ClockSampler:
    """nvidia-smi sampling at 20 ms in a child process for the whole run; samples are selected by
    wall-clock window afterwards (B200_PROFILING.md clocks line)."""
    Q = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
        "clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.lines = []
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _pump(self):
        for ln in self.p.stdout:
            self.lines.append((time.time(), ln))

    def window(self, t0, t1):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        rows = [ln.split(",") for (t, ln) in self.lines if t0 - 0.02 <= t <= t1 + 0.04]
        rows = [[x.strip() for x in r] for r in rows if len(r) >= 8]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples in window"]}
        sm = sorted(int(float(r[1])) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[4 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": int(float(rows[0][2])),
                "power_w_max": max(float(r[3]) for r in rows), "reasons": reasons, "samples": len(rows)}

    def close(self):
        if self.p is not None:
            self.p.terminate()


# ------------------------------------------------------------------------------------------------
# CPU leg: the oracle port of the reference on the host cores, on a bounded sample of the SAME workload
# ------------------------------------------------------------------------------------------------
def band_to_scipy_dia(band, M, w):
    """Host band[k, i] = R[i, i + k - w] (numpy, (2w+1) x M) -> scipy DIA matrix in fp64 (data[k, j] = R[j - off_k, j])."""
    import scipy.sparse
    data = np.zeros((2 * w + 1, M))
    for k in range(2 * w + 1):
        off = k - w
        if off >= 0:
            data[k, off:] = band[k, : M - off]
        else:
            data[k, : M + off] = band[k, -off:]
    return scipy.sparse.dia_matrix((data, np.arange(-w, w + 1)), shape=(M, M))


def build_sample(torch, Ms, w, seed, dev, use_library=True):
    """The bounded sample both arms run: the first Ms markers' worth of the benchmark workload, produced by the SAME
    generator and parameters as the M-marker problem (ldgen.banded_dia_device, N_ld = 4096, Bartlett taper, s, h2).
    Returns the device half band (for the GPU), Rused as scipy CSR in fp64 (what src/main.py:199-265 hands the
    reference solver for an .npz LD file), r, x0."""
    U, ldb, band, r, x0, _ = build_problem(torch, Ms, w, seed, dev, use_library=use_library)
    R = band_to_scipy_dia(band.cpu().numpy(), Ms, w).tocsr()
    del band
    return U, ldb, R, r, x0


def cpu_reference_run(R, r, M_full, iterations, probes, threads):
    """Time the CPU oracle (numpy/scipy restatement of src/sgvamp.py: same algorithm, same scipy CG, A = gamw R + gam2 I
    materialised every iteration as src/sgvamp.py:312 does) on the sample; the per-iteration time is scaled linearly in
    M (every per-iteration cost is O(M w)).  The sparse matvec is row-split over `threads` host threads (scipy's own
    csr_matvec is single-threaded); BLAS / OpenMP pools are limited to one thread while that pool is active, so the
    arm behaves the same with and without OMP_NUM_THREADS in the environment."""
    from threadpoolctl import threadpool_limits
    from oracle import sgvamp_oracle as orc
    Ms = R.shape[0]
    p = vamp_params(Ms)
    o = orc.VAMPOracle([n_gwas(Ms)], Ms, p["rho"], p["gamw"], p["gam1"], p["prior_vars"], p["prior_probs"])
    tm = {}
    with threadpool_limits(limits=1):
        t0 = time.perf_counter()
        out = o.infer([R], [r], iterations, cg_maxit=p["cg_maxit"], em_prior_maxit=p["em_prior_maxit"],
                      learn_gamw=True, lmmse_damp=False, prior_update="em", update_prior_from=1,
                      probe_fn=lambda k, it, M_: probes[k, it], timers=tm, threads=threads, materialise_A=True)
        dt = time.perf_counter() - t0
    its_per_s_sample = iterations / dt
    return dict(value=its_per_s_sample * Ms / M_full, sample_its_per_s=its_per_s_sample, sample_M=Ms,
                seconds=dt, cg_iters=[list(x[0]) for x in out["cg_iters"]], timers=tm, out=out)


def cpu_sample_text(res, w, its, threads):
    return ("oracle port of src/sgvamp.py (scipy CG semantics, A materialised per iteration, csr_matvec row-split over %d "
            "threads, BLAS pools limited to 1), %d VAMP iterations from it=0 on the first M=%d markers of the same "
            "generator (N_ld=%d, w=%d) (%.1f s), scaled by M_sample/M" % (threads, its, res["sample_M"], N_LD, w, res["seconds"]))


def gpu_sample_parity(sgvamp, U, ldb, w, r, ref_out, iterations, probes, device, stream):
    """Run the GPU path on the sample the CPU leg ran (same LD values, r, probes, parameters) and compare every
    iteration: the north-star gate (xhat rel-L2 <= 1e-4, scalars <= 1e-4 relative) measured inside the bench run."""
    Ms = len(r)
    p = vamp_params(Ms)
    v = sgvamp.VAMP(N=n_gwas(Ms), Nt=n_gwas(Ms), M=Ms, K=1, rho=p["rho"], gamw=p["gamw"], gam1=p["gam1"],
                    a=np.array([1.0]), prior_vars=p["prior_vars"], prior_probs=p["prior_probs"], out_dir=None,
                    out_name="parity", device=device, stream=stream)
    xs = v.infer(sgvamp.DeviceDSYM(U.data_ptr(), w, ldb, 0, keepalive=U), r, iterations, cg_maxit=p["cg_maxit"],
                 em_prior_maxit=p["em_prior_maxit"], learn_gamw=True, lmmse_damp=False, prior_update="em",
                 update_prior_from=1, probes=probes)
    xerr, serr, cg_equal = 0.0, 0.0, True
    for it in range(iterations):
        ref_x = ref_out["xhat1"][it]
        xerr = max(xerr, float(np.linalg.norm(xs[it].ravel() - ref_x) / np.linalg.norm(ref_x)))
        a_, b_ = np.array(v.history["rows"][it][0][1:7], dtype=np.float64), np.array(ref_out["rows"][it][0][1:7], dtype=np.float64)
        serr = max(serr, float(np.max(np.abs(a_ - b_) / np.abs(b_))))
        cg_equal = cg_equal and tuple(v.history["cg_iters"][it][0]) == tuple(ref_out["cg_iters"][it][0])
    cg = [list(v.history["cg_iters"][it][0]) for it in range(iterations)]
    v.close()
    return {"xhat_rel_l2_max": xerr, "scalar_rel_max": serr, "cg_iters_equal": bool(cg_equal), "cg_iters_gpu": cg,
            "cg_iters_ref": [list(x[0]) for x in ref_out["cg_iters"]], "iterations": iterations, "sample_M": Ms,
            "what": "GPU (C ABI, DSYM fused CG) vs CPU oracle on the cpu_baseline sample: same LD, r, probes; "
                    "scalars = gamw, gam1, gam2, alpha1, alpha2, lam"}


def main():
    a = parse()
    if a.config != "c5":
        import bench_configs
        if a.impl == "reference":
            return bench_configs.run_reference(a, sys.modules[__name__])
        return bench_configs.run_config(a, sys.modules[__name__])
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ncores = len(os.sched_getaffinity(0))
    workload = "banded LD M=%d w=%d (symmetric half band fp32 in HBM), K=1, L=2, EM prior, learn gamw, s=%.1f, cg_maxit=500" % (
        a.M, a.w, S_REG)

    # -------------------------------------------------------------------------------------------
    if a.impl == "reference":
        if rank != 0:
            return
        import torch
        dev = torch.device("cuda", local_rank) if torch.cuda.is_available() else torch.device("cpu")
        its = max(2, min(a.steps + a.warmup, 3))
        Ms = int(min(a.cpu_sample_M, a.M))
        _U, _ldb, Rs, rs_, _x0 = build_sample(torch, Ms, a.w, a.seed, dev, use_library=False)
        del _U
        res = cpu_reference_run(Rs, rs_, a.M, its, make_probes(its, Ms, a.seed), threads=ncores)
        line = {"impl": "reference", "metric": "VAMP iterations/s", "value": res["value"], "unit": "it/s",
                "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": 1000.0 / res["value"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload},
                "cpu_baseline": {"value": res["value"], "unit": "it/s", "cores": ncores, "kind": "port",
                                 "sample": cpu_sample_text(res, a.w, its, ncores),
                                 "sample_its_per_s": res["sample_its_per_s"], "cg_iters": res["cg_iters"],
                                 "timers_s": {k: round(x, 3) for k, x in res["timers"].items()}},
                "e2e": {"value": res["value"], "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # -------------------------------------------------------------------------------------------
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", ""):
            os.environ["NCCL_DEBUG"] = "WARN"      # keep NCCL's version banner off stdout: rank 0 prints ONE json line
        dist.init_process_group("nccl", device_id=dev)
    import build_native
    if rank == 0:
        build_native.build()
    if world > 1:
        dist.barrier()
    import sgvamp
    import shard as shd

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        t = torch.tensor(np.asarray(v, dtype=np.float64), device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.cpu().numpy()

    M, w = a.M, a.w
    iterations = a.warmup + a.steps
    shard = shd.TorchShard() if world > 1 else shd.SoloShard()
    bounds = shd.partition_rows(M, world)          # contiguous marker rows per GPU; halo of w entries from the neighbours
    lo, hi = bounds[rank]
    Ml = hi - lo
    ext = (w + 255) // 256 * 256 if rank > 0 else 0   # == sgv_dsym_extension (checked again by sgv_ld_adopt_dsym)
    U, ldb, band, r, x0, t_gen = build_problem(torch, M, w, a.seed, dev, lo, hi, ext, keep_full=not a.no_e2e)
    p = vamp_params(M)
    probes = make_probes(iterations, M, a.seed)     # global probes: every rank uses its own rows
    torch.cuda.synchronize()
    solver_stream = torch.cuda.Stream(device=dev)      # the library launches on this stream; torch events
    torch.cuda.set_stream(solver_stream)               # recorded below are recorded on the same stream
    stream = solver_stream.cuda_stream

    def new_solver():
        return sgvamp.VAMP(N=n_gwas(M), Nt=n_gwas(M), M=M, K=1, rho=p["rho"], gamw=p["gamw"], gam1=p["gam1"],
                           a=np.array([1.0]), prior_vars=p["prior_vars"], prior_probs=p["prior_probs"],
                           out_dir=None, out_name="bench", comm=None, device=local_rank, stream=stream,
                           shard=shard, shard_rows=bounds, halo=True)

    def run(v, R, n_it, hook=None, **kw):
        return v.infer(R, r, n_it, x0=None, cg_maxit=p["cg_maxit"], em_prior_maxit=p["em_prior_maxit"],
                       learn_gamw=p["learn_gamw"], lmmse_damp=p["lmmse_damp"], prior_update=p["prior_update"],
                       update_prior_from=p["update_prior_from"], probes=probes, iter_hook=hook,
                       gather_outputs=False, **kw)

    # ---- device-resident leg: LD already in HBM when the timed region starts ----
    sampler = ClockSampler(local_rank)
    dia = sgvamp.DeviceDSYM(U.data_ptr(), w, ldb, ext, keepalive=U)   # symmetric half band resident in HBM
    v0 = new_solver()
    run(v0, dia, 2)                                    # process-level warm-up (module load, allocations)
    v0.close()
    v = new_solver()                                   # fresh solver: prior / state as at program start
    events, launches, wall = {}, {}, {}

    def hook(it):
        if it == a.warmup:
            barrier()                                  # all ranks enter the timed region together
            v.handle.profile(True)
            launches["a"] = v.handle.launch_count()
            wall["a"] = time.time()
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        events[it] = e

    xs = run(v, dia, iterations, hook)
    barrier()
    wall["b"] = time.time()
    spmm_ms, spmm_launches = v.handle.profile_read()
    v.handle.profile(False)
    launches["b"] = v.handle.launch_count()
    clocks = sampler.window(wall["a"], wall["b"])
    ms_total = max_over_ranks(events[a.warmup].elapsed_time(events[iterations]))     # device time, max over ranks
    ms_from0 = max_over_ranks(events[0].elapsed_time(events[iterations]))
    hist = v.history
    passes = sum(hist["spmm_passes"][a.warmup:])
    info = v.handle.ld_info(0)
    value = a.steps / (ms_total / 1e3)
    # roofline of the dominant kernel (per GPU): algorithmic bytes of one 2-RHS pass over this rank's rows /
    # mean launch time (all SpMM launches of the timed region, early-exit launches included in the time but
    # not in the count of passes -> conservative); the slowest rank is reported
    bytes_pass = info["bytes_per_pass"]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    avg_ms = spmm_ms / max(passes, 1)
    achieved = -max_over_ranks(-(bytes_pass / (avg_ms * 1e-3) / 1e9))               # min over ranks
    avg_ms_max = max_over_ranks(avg_ms)
    spmm_share = max_over_ranks(spmm_ms) / ms_total
    # isolated kernel timing (back-to-back launches, inputs >> L2); single-rank hook
    iso_ms = v.handle.spmm_bench(0, 20) if world == 1 else None
    xl = x0[lo:hi]
    dots = sum_over_ranks([[float(np.dot(x.ravel(), xl)), float(np.dot(x.ravel(), x.ravel()))] for x in xs] +
                          [[float(np.dot(xl, xl)), 0.0]])
    aligns = [float(dots[i, 0] / max(np.sqrt(dots[i, 1] * dots[-1, 0]), 1e-300)) for i in range(iterations)]
    rows = [hist["rows"][i][0] for i in range(iterations)]
    if rank == 0:
        sys.stderr.write("host timers over all %d iterations (s): %s\n" % (iterations, {k: round(x, 4) for k, x in v.timers.items()}))
        sys.stderr.write("trajectory (it gamw gam1 gam2 alpha1 alpha2 lam | cg | align):\n")
        for i, rw in enumerate(rows):
            sys.stderr.write("  %2d %.4g %.4g %.4g %.4g %.4g %.4g | %s em=%d | %.4f\n" % (
                rw[0], rw[1], rw[2], rw[3], rw[4], rw[5], rw[6], hist["cg_iters"][i][0], hist["em_steps"][i], aligns[i]))

    # ---- end-to-end leg: host LD (scipy DIA arrays in pinned memory) -> VAMP.load_ld -> VAMP.infer -> host xhat ----
    e2e = None
    if not a.no_e2e:
        import scipy.sparse
        del U, dia
        glo = lo - ext
        if a.e2e_format == "csr":
            Rh, keep = band_to_host_csr(torch, band[:, ext:], M, w, lo, hi)
            h2d_ld = Rh.data.nbytes + Rh.indices.nbytes + (Ml + 1) * 8
        else:
            host, offsets, col0 = band_to_host_dia(torch, band, M, w, glo, hi)
            keep = host
            Rh = (scipy.sparse.dia_matrix((host.numpy(), offsets), shape=(M, M)) if world == 1
                  else sgvamp.DiaWindow(host.numpy(), offsets, col0, M))
            h2d_ld = (2 * w + 1) * (hi - glo) * 4            # upper diagonals + the lower ones for the symmetry check
        v2 = new_solver()
        v2.load_ld(0, Rh)                              # untimed warm-up of the upload path (staging buffers, first touch)
        barrier()
        os.environ["SGV_TIMING"] = "1"                 # upload phases to stderr (diagnosis of slow host links)
        t0 = time.perf_counter()
        v2.load_ld(0, Rh)
        torch.cuda.synchronize()
        t_up = time.perf_counter() - t0
        os.environ.pop("SGV_TIMING", None)
        t_inf = time.perf_counter()
        xs2 = run(v2, None, iterations, None, write_outputs=False)
        t_inf = time.perf_counter() - t_inf
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        if rank == 0:
            sys.stderr.write("e2e leg: upload %.4f s, infer %.4f s of which %s\n" % (t_up, t_inf, {k: round(x, 4) for k, x in v2.timers.items()}))
        h2d = sum_over_ranks([(h2d_ld + Ml * 8) / iterations + Ml])[0]
        d2h = M * 8 + 256 * world
        diff = max_over_ranks(max(np.linalg.norm(x1 - x2) / max(np.linalg.norm(x1), 1e-300) for x1, x2 in zip(xs, xs2)))
        e2e = {"value": iterations / dt, "unit": "it/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "what": "VAMP.load_ld(host LD of this rank's rows as scipy %s arrays, fp32, pinned) + VAMP.infer(r host) for %d "
                       "iterations from it=0: LD upload (symmetry verified on the device) + layout conversion + every "
                       "iteration's probe H2D and xhat D2H inside the timed region (wall clock between barriers, max over "
                       "ranks)" % (a.e2e_format.upper(), iterations),
               "seconds": dt, "ld_upload_seconds": max_over_ranks(t_up),
               "ld_upload_gbs": h2d_ld / max(t_up, 1e-9) / 1e9,
               "its_per_s_excluding_upload": iterations / max(dt - max_over_ranks(t_up), 1e-9),
               "its_per_s_resident_from_it0": iterations / (ms_from0 / 1e3),
               "note": "the upload moves %.1f GB per GPU over the host link once (half of it only to verify symmetry on the "
                       "device) and is amortised over %d iterations; the rest is VAMP.infer from it=0: per-run set-up (state "
                       "reset, pinned log / probe buffers) + the loop, which runs at the resident leg's rate from it=0 (the "
                       "timed window of `value` starts at it=%d, after the long CG solves of the first iterations)" % (
                           h2d_ld / 1e9, iterations, a.warmup),
               "infer_seconds": t_inf, "infer_setup_seconds": v2.timers.get("setup"), "infer_loop_seconds": v2.timers.get("loop"),
               "max_rel_diff_vs_resident": diff,
               "layout": v2.handle.ld_info(0)["layout"], "host_format": a.e2e_format}
        if a.e2e_format == "dia":
            # the same leg with the container declared symmetric (what the reference implicitly assumes: it never checks):
            # only the upper diagonals travel
            v2.close()
            v2 = new_solver()                          # a fresh solver: the first run's learned prior must not carry over
            v2.load_ld(0, Rh, assume_symmetric=True)   # untimed warm-up of the upload path, as above
            barrier()
            os.environ["SGV_TIMING"] = "1"
            t0s = time.perf_counter()
            v2.load_ld(0, Rh, assume_symmetric=True)
            torch.cuda.synchronize()
            t_up_s = time.perf_counter() - t0s
            os.environ.pop("SGV_TIMING", None)
            xs3 = run(v2, None, iterations, None, write_outputs=False)
            barrier()
            dt_s = max_over_ranks(time.perf_counter() - t0s)
            diff_s = max_over_ranks(max(np.linalg.norm(x1 - x3) / max(np.linalg.norm(x1), 1e-300) for x1, x3 in zip(xs, xs3)))
            e2e["declared_symmetric"] = {"value": iterations / dt_s, "unit": "it/s", "seconds": dt_s,
                                         "ld_upload_seconds": max_over_ranks(t_up_s),
                                         "h2d_bytes_per_step": int(sum_over_ranks([((w + 1) * (hi - glo) * 4 + Ml * 8) / iterations + Ml])[0]),
                                         "max_rel_diff_vs_resident": diff_s,
                                         "what": "VAMP.load_ld(..., assume_symmetric=True) + VAMP.infer: the lower diagonals are "
                                                 "not shipped for verification"}
        v2.close()
        del Rh, keep

    cpu, parity = None, None
    if not a.no_cpu_baseline and world == 1:
        its_c = 2
        Ms = int(min(a.cpu_sample_M, M))
        Us, ldbs, Rs, rs_, _x0s = build_sample(torch, Ms, w, a.seed, dev)
        probes_s = make_probes(its_c, Ms, a.seed)
        res = cpu_reference_run(Rs, rs_, M, its_c, probes_s, threads=ncores)
        cpu = {"value": res["value"], "unit": "it/s", "cores": ncores, "kind": "port",
               "sample": cpu_sample_text(res, w, its_c, ncores), "sample_its_per_s": res["sample_its_per_s"]}
        parity = gpu_sample_parity(sgvamp, Us, ldbs, w, rs_, res["out"], its_c, probes_s, local_rank, stream)
        del Us, Rs
    sampler.close()
    traffic = None
    try:   # dram bytes per launch of the dominant kernel from the committed ncu --set full capture (1 GPU, M=1M, w=500)
        if world == 1 and M == 1_000_000 and w == 500:
            traffic = json.load(open(os.path.join(REPO, "profiles", "traffic.json")))["spmm_bytes_per_pass"]
    except Exception:
        pass

    line = {
        "metric": "VAMP iterations/s", "value": value, "unit": "it/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "M": M, "half_bandwidth": w, "layout": info["layout"],
                   "partition": "%d contiguous row shards, halo of %d vector entries read from the neighbours' HBM over NVLink, "
                                "scalar reductions exchanged in-kernel through peer memory" % (world, w) if world > 1 else "single GPU",
                   "nnz_stored_per_gpu": info["nnz_stored"],
                   "l2_policy": "inputs (%.2f GB band per GPU) larger than L2" % (info["nnz_stored"] * 4 / 1e9),
                   "timed_iterations": "VAMP iterations %d..%d of one trajectory" % (a.warmup, iterations - 1),
                   "cg_iters_timed": [list(hist["cg_iters"][i][0]) for i in range(a.warmup, iterations)],
                   "spmm_passes_timed": passes, "its_per_s_from_it0": iterations / (ms_from0 / 1e3),
                   "alignment_with_truth": aligns[-1], "gen_seconds": t_gen},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "k_dsym_persist (2-RHS symmetric half-band SpMM over persistent row ranges with the whole CG step fused: direction / "
                               "residual / solution updates while the window is staged, q and the four dot products in the epilogue; one "
                               "cooperative launch per solve, avg_launch_ms = solve kernel time / CG steps)",
                     "per": "GPU (slowest rank)", "bytes_per_launch": bytes_pass, "avg_launch_ms": avg_ms_max,
                     "launches_timed": spmm_launches, "isolated_launch_ms": iso_ms,
                     "isolated_gbs": (bytes_pass / (iso_ms * 1e-3) / 1e9) if iso_ms else None,
                     "peak_source": peak_src, "spmm_share_of_step": spmm_share},
        "cpu_baseline": cpu, "parity": parity, "e2e": e2e, "gpu_launches": int(sum_over_ranks([launches["b"] - launches["a"]])[0]),
        "clocks": clocks,
    }
    if rank == 0:
        print(json.dumps(line))
    v.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
