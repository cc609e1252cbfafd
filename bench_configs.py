"""bench.py --config c1|c2|c3|c4: the other shapes BASELINE.json names (dense and block-diagonal LD), one GPU.

    c1  M=10k dense, N=10k, K=1, L=2, 10 iterations                      (the reference's CPU-runnable case)
    c2  M=50k dense, N=100k, K=1, L=4, cg-maxit 50, learned gamw
    c3  M=300k block-diagonal LD (LD blocks of 500..3500 markers), K=1, L=2, s=0.1
    c4  K=3 cohorts x M=100k dense (N = 200k/300k/250k, shared effects), EM prior update

Inputs follow SURVEY 8(d): the dense shapes use the reference's own recipe (simulation/sim_gen_phen_mult.py:28-55:
X ~ Binomial(2, 0.4), column-standardised, sparse beta, y = X beta + noise, r = X^T y / sqrt(N), R = X^T X / N), with
X^T X accumulated chunk-wise on the GPU (torch as a data-generation utility; genotypes are small integers, so the
TF32 products and the fp32 accumulation are exact); the block-diagonal shape draws every LD block from the thresholded
latent-Gaussian haplotype generator of ldgen.py (N_ld = 4096).  Same JSON contract as the default (c5) line.
"""
import json
import os
import sys
import time

import numpy as np

H2, LAM_TRUE = 0.5, 0.01

CONFIGS = {
    "c1": dict(kind="dense", M=10_000, N=[10_000], L=2, s=0.0, cg_maxit=500, iterations=10, sample_M=3_000,
               text="M=10k dense LD, N=10k, K=1, L=2 (spike+slab), EM prior, learn gamw, s=0"),
    "c2": dict(kind="dense", M=50_000, N=[100_000], L=4, s=0.0, cg_maxit=50, iterations=10, sample_M=4_000,
               text="M=50k dense LD, N=100k, K=1, L=4 mixture, cg-maxit=50, EM prior, learned gamw, s=0"),
    "c3": dict(kind="blockdiag", M=300_000, N=[600_000], L=2, s=0.1, cg_maxit=500, iterations=10, sample_M=24_000, rho=0.2,
               text="M=300k block-diagonal LD (LD blocks of 500-3500 markers, dense panels), K=1, L=2, EM prior, learn gamw, s=0.1"),
    "c4": dict(kind="dense", M=100_000, N=[200_000, 300_000, 250_000], L=2, s=0.0, cg_maxit=500, iterations=10, sample_M=2_500,
               text="K=3 cohorts x M=100k dense LD (N=200k/300k/250k, shared effects), L=2, EM prior update, learn gamw, s=0"),
}


def prior_for(M, L):
    cm = max(1, int(M * LAM_TRUE))
    v = H2 / cm
    if L == 2:
        return [0.0, v], [0.99, 0.01]
    return [0.0, 0.1 * v, v, 10 * v], [0.97, 0.01, 0.01, 0.01]


def causal_beta(M, seed):
    rng = np.random.default_rng(seed + 99)
    cm = max(1, int(M * LAM_TRUE))
    beta = np.zeros(M)
    beta[rng.choice(M, cm, replace=False)] = rng.normal(0, np.sqrt(H2 / cm), cm)
    return beta


def gen_dense_cohort(torch, M, N, beta, seed, dev, out=None, chunk=8192, cols=None):
    """R = X^T X / N (fp32, M x M on the device, exactly symmetric) and r = X^T y / sqrt(N) (host fp64) of the reference
    recipe, in two passes over seeded row chunks of X (pass 1: column sums and the raw Gram matrix; pass 2: y and r).
    cols = (lo, hi): only the column panel R[:, lo:hi] (M x roundup(hi-lo, 4), what one rank of the dense rows partition
    holds; the same X on every rank), standardised with a formula that is symmetric in (i, j) to the last bit."""
    torch.backends.cuda.matmul.allow_tf32 = True        # entries 0/1/2: products and fp32 sums are exact
    if cols is not None:
        return _gen_dense_colpanel(torch, M, N, beta, seed, dev, chunk, cols)
    G = out if out is not None else torch.empty((M, M), device=dev, dtype=torch.float32)
    G.zero_()
    s1 = torch.zeros(M, device=dev, dtype=torch.float64)

    def chunk_rows(c0):
        g = torch.Generator(device=dev)
        g.manual_seed(seed * 7919 + c0)
        n = min(chunk, N - c0)
        u = torch.rand((n, M), generator=g, device=dev, dtype=torch.float32)
        return (u < 0.4).to(torch.float32) + (torch.rand((n, M), generator=g, device=dev, dtype=torch.float32) < 0.4).to(torch.float32)

    for c0 in range(0, N, chunk):
        X = chunk_rows(c0)
        G.addmm_(X.t(), X)
        s1 += X.sum(dim=0, dtype=torch.float64)
        del X
    mu = s1 / N
    diag = torch.diagonal(G).to(torch.float64)
    sd = torch.sqrt(torch.clamp(diag / N - mu * mu, min=1e-12))
    bt = torch.from_numpy(beta).to(dev)
    r = torch.zeros(M, device=dev, dtype=torch.float64)
    gn = torch.Generator(device=dev)
    gn.manual_seed(seed * 31 + 5)
    for c0 in range(0, N, chunk):
        X = chunk_rows(c0).to(torch.float64)
        Xs = (X - mu[None, :]) / sd[None, :]
        y = Xs @ bt + float(np.sqrt(1 - H2)) * torch.randn((Xs.shape[0],), generator=gn, device=dev, dtype=torch.float64)
        r += Xs.t() @ y
        del X, Xs
    r /= float(np.sqrt(N))
    # standardise the Gram matrix in place, row block by row block: R = (G - N mu mu^T) / (N sd sd^T)
    mu32, sd32 = mu.to(torch.float32), sd.to(torch.float32)
    for i0 in range(0, M, 4096):
        i1 = min(M, i0 + 4096)
        blk = G[i0:i1]
        blk -= float(N) * mu32[i0:i1, None] * mu32[None, :]
        blk /= (float(N) * sd32[i0:i1, None] * sd32[None, :])
    # fp32 rounding differs between the two triangles by an ulp: mirror the upper one (LD must be symmetric)
    for i0 in range(0, M, 4096):
        i1 = min(M, i0 + 4096)
        G[i0:i1, :i0] = G[:i0, i0:i1].t()
        d = G[i0:i1, i0:i1]
        d.copy_(torch.triu(d) + torch.triu(d, 1).t())
    torch.diagonal(G).fill_(1.0)
    if dev.type == "cuda":
        torch.cuda.synchronize()
    return G, r.cpu().numpy()


def _dense_chunk(torch, M, N, seed, dev, chunk, c0):
    g = torch.Generator(device=dev)
    g.manual_seed(seed * 7919 + c0)
    n = min(chunk, N - c0)
    u = torch.rand((n, M), generator=g, device=dev, dtype=torch.float32)
    return (u < 0.4).to(torch.float32) + (torch.rand((n, M), generator=g, device=dev, dtype=torch.float32) < 0.4).to(torch.float32)


def _gen_dense_colpanel(torch, M, N, beta, seed, dev, chunk, cols):
    lo, hi = cols
    Ml = hi - lo
    ldd = (Ml + 3) // 4 * 4
    P = torch.zeros((M, ldd), device=dev, dtype=torch.float32)
    s1 = torch.zeros(M, device=dev, dtype=torch.float64)
    s2 = torch.zeros(M, device=dev, dtype=torch.float64)
    for c0 in range(0, N, chunk):
        X = _dense_chunk(torch, M, N, seed, dev, chunk, c0)
        P[:, :Ml].addmm_(X.t(), X[:, lo:hi])
        s1 += X.sum(dim=0, dtype=torch.float64)
        s2 += (X * X).sum(dim=0, dtype=torch.float64)
        del X
    mu = s1 / N
    sd = torch.sqrt(torch.clamp(s2 / N - mu * mu, min=1e-12))
    bt = torch.from_numpy(beta).to(dev)
    r = torch.zeros(M, device=dev, dtype=torch.float64)
    gn = torch.Generator(device=dev)
    gn.manual_seed(seed * 31 + 5)
    for c0 in range(0, N, chunk):
        X = _dense_chunk(torch, M, N, seed, dev, chunk, c0).to(torch.float64)
        Xs = (X - mu[None, :]) / sd[None, :]
        y = Xs @ bt + float(np.sqrt(1 - H2)) * torch.randn((Xs.shape[0],), generator=gn, device=dev, dtype=torch.float64)
        r += Xs.t() @ y
        del X, Xs
    r /= float(np.sqrt(N))
    mu32, sd32 = mu.to(torch.float32), sd.to(torch.float32)
    for i0 in range(0, M, 4096):
        i1 = min(M, i0 + 4096)
        blk = P[i0:i1, :Ml]
        blk -= float(N) * (mu32[i0:i1, None] * mu32[None, lo:hi])          # products commute: R[i][j] == R[j][i] bit for bit
        blk /= (float(N) * (sd32[i0:i1, None] * sd32[None, lo:hi]))
    idx = torch.arange(lo, hi, device=dev)
    P[idx, idx - lo] = 1.0
    if dev.type == "cuda":
        torch.cuda.synchronize()
    return P, r.cpu().numpy()


def block_sizes(M, seed):
    rng = np.random.default_rng(seed)
    sizes, left = [], M
    while left > 0:
        b = int(min(left, rng.integers(500, 3500)))
        if left - b < 500 and left - b > 0:
            b = left
        sizes.append(b)
        left -= b
    return sizes


def gen_blockdiag(torch, M, seed, dev, s, N_gwas, N_ld=4096, row_lo=0, row_hi=None):
    """Block-diagonal LD: panels (one dense R_b = X_b^T X_b per LD block, Rused applied), block starts / offsets / lds,
    r and x0.  With a row range (a rank of a block partition: boundaries are block boundaries) only the blocks inside it
    are generated; starts are then local.  Every block depends only on (seed, block index)."""
    import ldgen
    row_hi = M if row_hi is None else row_hi
    all_sizes = block_sizes(M, seed)
    all_starts = np.concatenate([[0], np.cumsum(all_sizes)]).astype(np.int64)
    mine = [b for b in range(len(all_sizes)) if all_starts[b] >= row_lo and all_starts[b + 1] <= row_hi]
    assert mine and all_starts[mine[0]] == row_lo and all_starts[mine[-1] + 1] == row_hi, "row range must consist of whole blocks"
    sizes = [all_sizes[b] for b in mine]
    starts = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    lds = np.array([(m + 3) // 4 * 4 for m in sizes], dtype=np.int32)
    offs = np.concatenate([[0], np.cumsum(np.array(sizes, dtype=np.int64) * lds)])[:-1].astype(np.int64)
    total = int(np.sum(np.array(sizes, dtype=np.int64) * lds))
    P = torch.zeros(total, device=dev, dtype=torch.float32)
    beta = causal_beta(M, seed)
    x0 = beta * np.sqrt(N_gwas)
    x0d = torch.from_numpy(x0[row_lo:row_hi]).to(dev)
    r = torch.zeros(row_hi - row_lo, device=dev, dtype=torch.float64)
    g = torch.Generator(device=dev)
    for b, m in enumerate(sizes):
        lo = int(starts[b])
        gb = mine[b]                                                        # global block index
        glo = int(all_starts[gb]) + gb * 4096                               # offset: blocks independent
        g.manual_seed(seed * 31 + 17 + gb * 1009)
        X = ldgen.genotypes_device(torch, N_ld, glo, glo + m, seed, dev)
        Rb = X.t() @ X
        Rb = torch.triu(Rb) + torch.triu(Rb, 1).t()
        Rb.fill_diagonal_(1.0)
        Rb = Rb * (1.0 - s)
        Rb.diagonal().add_(s)
        ld = int(lds[b])
        P[int(offs[b]): int(offs[b]) + m * ld].view(m, ld)[:, :m] = Rb
        z = torch.randn((N_ld,), generator=g, device=dev, dtype=torch.float64)
        zi = torch.randn((m,), generator=g, device=dev, dtype=torch.float64)
        noise = float(np.sqrt(1 - s)) * (X.to(torch.float64).t() @ z) + float(np.sqrt(s)) * zi   # ~ N(0, Rused_b)
        r[lo:lo + m] = Rb.to(torch.float64) @ x0d[lo:lo + m] + float(np.sqrt(1 - H2)) * noise
        del X, Rb
    if dev.type == "cuda":
        torch.cuda.synchronize()
    return P, starts, offs, lds, r.cpu().numpy(), x0


def cpu_leg(torch, cfg, bench, seed, dev, ncores):
    """The oracle port of the reference on a bounded sample of the configuration (same generator, reduced M), timed on
    the host cores.  Returns the cpu_baseline object, the oracle's records and the sample."""
    from threadpoolctl import threadpool_limits
    from oracle import sgvamp_oracle as orc
    M, Ns, L, s, K = cfg["M"], cfg["N"], cfg["L"], cfg["s"], len(cfg["N"])
    Ms = cfg["sample_M"]
    its_c = 3
    betas = causal_beta(Ms, seed)
    Rh, rh, Nss = [], [], []
    if cfg["kind"] == "dense":
        for k, N in enumerate(Ns):
            Nk = max(Ms, int(round(N * Ms / M)))
            G, r = gen_dense_cohort(torch, Ms, Nk, betas, seed + k, dev)
            Rh.append(G.cpu().numpy().astype(np.float64))
            rh.append(r)
            Nss.append(Nk)
        scale = (Ms / M) ** 2
    else:
        import scipy.sparse
        P, st_, of_, ld_, r, _x0 = gen_blockdiag(torch, Ms, seed, dev, s, 2 * Ms)
        Pc = P.cpu().numpy()
        blocks = [scipy.sparse.csr_matrix(Pc[int(of_[b]): int(of_[b]) + int(st_[b + 1] - st_[b]) * int(ld_[b])].reshape(
            int(st_[b + 1] - st_[b]), int(ld_[b]))[:, : int(st_[b + 1] - st_[b])].astype(np.float64)) for b in range(len(ld_))]
        Rh.append(scipy.sparse.block_diag(blocks, format="csr"))
        rh.append(r)
        Nss.append(2 * Ms)
        scale = Ms / M
    pvs, pps = prior_for(Ms, L)
    pr_s = np.stack([bench.make_probes(its_c, Ms, seed + k)[0] for k in range(K)])
    o = orc.VAMPOracle(Nss, Ms, cfg.get("rho", 0.5), 2.0, 1e-6, pvs, pps)
    dense = cfg["kind"] == "dense"
    with threadpool_limits(limits=ncores if dense else 1):
        t0 = time.perf_counter()
        ref = o.infer(Rh, rh, its_c, cg_maxit=cfg["cg_maxit"], em_prior_maxit=100, learn_gamw=True, lmmse_damp=False,
                      prior_update="em", update_prior_from=1, probe_fn=lambda k, it, M_: pr_s[k, it], materialise_A=True,
                      threads=1 if dense else ncores)
        dt = time.perf_counter() - t0
    cpu = {"value": its_c / dt * scale, "unit": "it/s", "cores": ncores, "kind": "port",
           "sample": "oracle port of src/sgvamp.py (scipy CG semantics, A = gamw R + gam2 I materialised per iteration as "
                     "src/sgvamp.py:312 does, %s), %d VAMP iterations from it=0 at M=%d from the same generator (%.1f s), "
                     "scaled by %s" % ("numpy/BLAS matvec on %d threads" % ncores if dense else
                                       "csr_matvec row-split over %d threads, BLAS pools limited to 1" % ncores, its_c, Ms, dt,
                                       "(M_sample/M)^2" if dense else "M_sample/M"),
           "sample_its_per_s": its_c / dt}
    return cpu, ref, (Rh, rh, Nss, Ms, pr_s, its_c)


def run_reference(a, bench):
    """--impl reference for c1..c4: the CPU leg alone (rank 0), same JSON shape as the c5 reference line."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    cfg = CONFIGS[a.config]
    ncores = len(os.sched_getaffinity(0))
    dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    cpu, _ref, _sample = cpu_leg(torch, cfg, bench, a.seed, dev, ncores)
    line = {"impl": "reference", "metric": "VAMP iterations/s", "value": cpu["value"], "unit": "it/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1000.0 / cpu["value"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s: %s, cg_maxit=%d" % (a.config, cfg["text"], cfg["cg_maxit"])}, "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_rank_per_cohort(a, bench):
    """c4 under torchrun with WORLD_SIZE == K: the reference's own deployment shape, one rank (here: one GPU) per cohort
    (src/main.py:85,173-174); r1 / gam1 of every cohort are exchanged once per iteration (src/sgvamp.py:228-233) through
    shard.TorchComm over NCCL.  Each rank generates and holds only its cohort's LD (40 GB at M = 100k)."""
    import torch
    import torch.distributed as dist
    import build_native
    cfg = CONFIGS[a.config]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(local_rank)
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", ""):
        os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        build_native.build()
    dist.barrier()
    import sgvamp
    import shard as shd
    M, Ns, L, K = cfg["M"], cfg["N"], cfg["L"], len(cfg["N"])
    assert world == K
    iterations = a.warmup + a.steps
    beta = causal_beta(M, a.seed)
    t0 = time.time()
    G, r = gen_dense_cohort(torch, M, Ns[rank], beta, a.seed + rank, dev)
    t_gen = time.time() - t0
    probes = bench.make_probes(iterations, M, a.seed + rank)[0]
    solver_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(solver_stream)
    pv, pp = prior_for(M, L)

    def new_solver():
        return sgvamp.VAMP(N=Ns[rank], Nt=float(sum(Ns)), M=M, K=K, rho=cfg.get("rho", 0.5), gamw=2.0, gam1=1e-6,
                           a=np.array(Ns) / float(sum(Ns)), prior_vars=pv, prior_probs=pp, out_dir=None, out_name="bench",
                           comm=shd.TorchComm(), device=local_rank, stream=solver_stream.cuda_stream)

    def run(v, n_it, hook=None):
        return v.infer(sgvamp.DeviceDense(G.data_ptr(), M, keepalive=G), r, n_it, cg_maxit=cfg["cg_maxit"], em_prior_maxit=100,
                       learn_gamw=True, lmmse_damp=False, prior_update="em", update_prior_from=1,
                       probes=lambda k, it, M_: probes[it], iter_hook=hook, s=0.0)

    v0 = new_solver()
    run(v0, 2)
    v0.close()
    v = new_solver()
    events, launches = {}, {}

    def hook(it):
        if it == a.warmup:
            torch.cuda.synchronize()
            dist.barrier()
            v.handle.profile(True)
            launches["a"] = v.handle.launch_count()
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        events[it] = e

    xs = run(v, iterations, hook)
    torch.cuda.synchronize()
    dist.barrier()
    spmm_ms, _n = v.handle.profile_read()
    launches["b"] = v.handle.launch_count()
    t = torch.tensor([events[a.warmup].elapsed_time(events[iterations]), float(launches["b"] - launches["a"])], device=dev, dtype=torch.float64)
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    ms_total = float(tmax[0].item())
    passes = sum(v.history["spmm_passes"][a.warmup:])
    info = v.handle.ld_info(rank)
    bytes_pass = 2.0 * info["nnz_stored"] + 32.0 * M
    avg_ms = spmm_ms / max(passes, 1)
    x0 = beta * np.sqrt(Ns[0])
    align = float(np.dot(xs[-1].ravel(), x0) / (np.linalg.norm(xs[-1]) * np.linalg.norm(x0)))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(bench.REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    if rank == 0:
        print(json.dumps({
            "metric": "VAMP iterations/s", "value": a.steps / (ms_total / 1e3), "unit": "it/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s: %s, cg_maxit=%d" % (a.config, cfg["text"], cfg["cg_maxit"]), "M": M, "K": K,
                       "partition": "one cohort per GPU (rank = cohort, the reference's deployment shape); r1 / gam1 exchanged per "
                                    "iteration over NCCL (shard.TorchComm)", "layout": info["layout"],
                       "cg_iters_timed_rank0": [list(v.history["cg_iters"][i][rank]) for i in range(a.warmup, iterations)],
                       "alignment_with_truth": align, "gen_seconds": t_gen},
            "roofline": {"bound": "hbm", "achieved": bytes_pass / (avg_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": bytes_pass / (avg_ms * 1e-3) / 1e9 / peak, "traffic": None, "per": "GPU (rank 0)",
                         "kernel": "k_spmm_psym + k_psym_finish", "bytes_per_launch": bytes_pass, "avg_launch_ms": avg_ms},
            "cpu_baseline": None, "parity": None, "e2e": None, "gpu_launches": int(t[1].item())}))
    v.close()
    dist.barrier()
    dist.destroy_process_group()


def run_config(a, bench):
    if a.config == "c4" and int(os.environ.get("WORLD_SIZE", "1")) == len(CONFIGS["c4"]["N"]):
        return run_rank_per_cohort(a, bench)
    import torch
    import build_native
    build_native.build()
    import sgvamp
    cfg = CONFIGS[a.config]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # c3: LD blocks sharded over the GPUs, scalar-only exchange; dense shapes: every cohort partitioned by rows over all
    # GPUs, the vector pair gathered from the peers before each product (SGV_DENSE_ROWS=0: rank 0 alone runs the shape)
    rows_part = world > 1 and cfg["kind"] == "dense" and os.environ.get("SGV_DENSE_ROWS", "1") != "0"
    sharded = world > 1 and (cfg["kind"] == "blockdiag" or rows_part)
    if world > 1 and not sharded:
        # the dense shapes are measured on one GPU; further ranks have nothing to do
        if rank != 0:
            return
        local_rank = 0
    ncores = len(os.sched_getaffinity(0))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(local_rank)
    import shard as shd
    shard = shd.SoloShard()
    if sharded:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", ""):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
        shard = shd.TorchShard()

    def max_over_ranks(x):
        if not sharded:
            return float(x)
        t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor(np.atleast_1d(np.asarray(x, dtype=np.float64)), device=dev)
        if sharded:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.cpu().numpy()
    M, Ns, L, s, K = cfg["M"], cfg["N"], cfg["L"], cfg["s"], len(cfg["N"])
    Nt = float(sum(Ns))
    pv, pp = prior_for(M, L)
    iterations = a.warmup + a.steps
    seed = a.seed
    t0 = time.time()
    keep, Rs, rs_, x0 = [], [], [], None
    beta = causal_beta(M, seed)
    bounds = None
    if cfg["kind"] == "dense" and rows_part:
        bounds = shd.partition_rows(M, world)
        lo, hi = bounds[rank]
        for k, N in enumerate(Ns):
            P, r = gen_dense_cohort(torch, M, N, beta, seed + k, dev, cols=(lo, hi))
            keep.append(P)
            Rs.append(sgvamp.DeviceDenseCols(P.data_ptr(), P.shape[1], keepalive=P))
            rs_.append(r[lo:hi])
        x0 = beta * np.sqrt(Ns[0])
    elif cfg["kind"] == "dense":
        for k, N in enumerate(Ns):
            G, r = gen_dense_cohort(torch, M, N, beta, seed + k, dev)
            keep.append(G)
            Rs.append(sgvamp.DeviceDense(G.data_ptr(), M, keepalive=G))
            rs_.append(r)
        x0 = beta * np.sqrt(Ns[0])
    else:
        lo, hi = 0, M
        if sharded:
            gstarts = np.concatenate([[0], np.cumsum(block_sizes(M, seed))])
            bounds = shd.partition_blocks(gstarts, world)                   # whole LD blocks per GPU, balanced by sum m_b^2
            lo, hi = bounds[rank]
        P, starts, offs, lds, r, x0 = gen_blockdiag(torch, M, seed, dev, s, Ns[0], row_lo=lo, row_hi=hi)
        keep.append(P)
        Rs.append(sgvamp.DeviceBlockDiag(P.data_ptr(), starts, offs, lds, keepalive=P))
        rs_.append(r)
    t_gen = time.time() - t0
    probes = np.stack([bench.make_probes(iterations, M, seed + k)[0] for k in range(K)])
    solver_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(solver_stream)
    stream = solver_stream.cuda_stream

    def new_solver(Mx=M, Nsx=Ns, solo=False):
        kw = dict(shard=shard, shard_rows=bounds, halo="rows" if rows_part else False) if (sharded and not solo) else {}
        return sgvamp.VAMP(N=Nsx if K > 1 else Nsx[0], Nt=float(sum(Nsx)), M=Mx, K=K, rho=cfg.get("rho", 0.5), gamw=2.0, gam1=1e-6,
                           a=np.array(Nsx) / float(sum(Nsx)), prior_vars=prior_for(Mx, L)[0], prior_probs=prior_for(Mx, L)[1],
                           out_dir=None, out_name="bench", device=local_rank, stream=stream, **kw)

    def run(v, R, r, n_it, pr, hook=None):
        return v.infer(R if K > 1 else R[0], list(r) if K > 1 else r[0], n_it, cg_maxit=cfg["cg_maxit"], em_prior_maxit=100,
                       learn_gamw=True, lmmse_damp=False, prior_update="em", update_prior_from=1, probes=pr, iter_hook=hook,
                       s=0.0, gather_outputs=False)

    sampler = bench.ClockSampler(local_rank)
    v0 = new_solver()
    run(v0, Rs, rs_, 2, probes)                                       # process-level warm-up
    v0.close()
    v = new_solver()
    events, launches, wall = {}, {}, {}

    def hook(it):
        if it == a.warmup:
            torch.cuda.synchronize()
            if sharded:
                dist.barrier()
            v.handle.profile(True)
            launches["a"] = v.handle.launch_count()
            wall["a"] = time.time()
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        events[it] = e

    xs = run(v, Rs, rs_, iterations, probes, hook)
    torch.cuda.synchronize()
    if sharded:
        dist.barrier()
    wall["b"] = time.time()
    spmm_ms, spmm_launches = v.handle.profile_read()
    v.handle.profile(False)
    launches["b"] = v.handle.launch_count()
    clocks = sampler.window(wall["a"], wall["b"])
    ms_total = max_over_ranks(events[a.warmup].elapsed_time(events[iterations]))
    hist = v.history
    passes = sum(hist["spmm_passes"][a.warmup:])
    infos = [v.handle.ld_info(k) for k in range(K)]
    # algorithmic bytes of one 2-RHS pass over a SYMMETRIC dense store: the upper triangle once (2 bytes per matrix
    # entry on average) + vector pair in and out
    Ml = v.Ml
    if rows_part:      # the rank's column panel once (no symmetry to use across ranks) + the gathered pair + pair in / out
        bytes_pass = float(np.mean([4.0 * i["nnz_stored"] + 16.0 * M + 32.0 * Ml for i in infos]))
    else:
        bytes_pass = float(np.mean([2.0 * i["nnz_stored"] + 32.0 * Ml for i in infos]))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(bench.REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    avg_ms = spmm_ms / max(passes, 1)
    achieved = -max_over_ranks(-(bytes_pass / (avg_ms * 1e-3) / 1e9))     # the slowest rank
    avg_ms = max_over_ranks(avg_ms)
    iso_ms = v.handle.spmm_bench(0, 20) if not sharded else None
    xl = x0[v.lo:v.hi]
    dots = sum_over_ranks([[float(np.dot(x.ravel(), xl)), float(np.dot(x.ravel(), x.ravel()))] for x in xs] + [[float(np.dot(xl, xl)), 0.0]])
    dots = dots.reshape(-1, 2)
    aligns = [float(dots[i, 0] / max(np.sqrt(dots[i, 1] * dots[-1, 0]), 1e-300)) for i in range(iterations)]
    nlaunch = int(sum_over_ranks([launches["b"] - launches["a"]])[0])
    nnz_all = [int(x) for x in sum_over_ranks([i["nnz_stored"] for i in infos])] if sharded else [i["nnz_stored"] for i in infos]
    if rank != 0:
        v.close()
        dist.barrier()
        dist.destroy_process_group()
        return
    sys.stderr.write("trajectory (it gamw gam1 gam2 alpha1 alpha2 lam | cg | align):\n")
    for i in range(iterations):
        rw = hist["rows"][i][0]
        sys.stderr.write("  %2d %.4g %.4g %.4g %.4g %.4g %.4g | %s em=%d | %.4f\n" % (
            rw[0], rw[1], rw[2], rw[3], rw[4], rw[5], rw[6], [hist["cg_iters"][i][k] for k in range(K)], hist["em_steps"][i], aligns[i]))
    value = a.steps / (ms_total / 1e3)
    layout = infos[0]["layout"]
    v.close()
    if sharded:
        torch.cuda.synchronize()

    # ---- end-to-end leg: host LD (fp32, pinned where it fits) -> VAMP.load_ld -> VAMP.infer -> host xhat
    e2e = None
    host_bytes = sum(4.0 * M * M for _ in range(K)) if cfg["kind"] == "dense" else 8.0 * infos[0]["nnz_stored"]
    if sharded:
        a.no_e2e = a.no_cpu_baseline = True        # rank 0 only prints; the one-GPU run of the shape carries those legs
    if not a.no_e2e and host_bytes < 48e9:
        import scipy.sparse
        hostR = []
        if cfg["kind"] == "dense":
            for G in keep:
                h_ = torch.empty((M, M), dtype=torch.float32, pin_memory=True)
                h_.copy_(G)
                hostR.append(h_.numpy())
        else:
            blocks = []
            Pc = keep[0].cpu().numpy()
            for b in range(len(lds)):
                m = int(starts[b + 1] - starts[b])
                blocks.append(scipy.sparse.csr_matrix(Pc[int(offs[b]): int(offs[b]) + m * int(lds[b])].reshape(m, int(lds[b]))[:, :m]))
            Rh = scipy.sparse.block_diag(blocks, format="csr")
            Rh.sort_indices()
            hostR.append(Rh)
            del Pc, blocks
        torch.cuda.synchronize()
        v2 = new_solver()
        for k in range(K):
            v2.load_ld(k, hostR[k])                                   # untimed warm-up of the upload path
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(K):
            v2.load_ld(k, hostR[k])
        torch.cuda.synchronize()
        t_up = time.perf_counter() - t0
        xs2 = v2.infer([None] * K if K > 1 else None, list(rs_) if K > 1 else rs_[0], iterations, cg_maxit=cfg["cg_maxit"],
                       em_prior_maxit=100, learn_gamw=True, lmmse_damp=False, prior_update="em", update_prior_from=1,
                       probes=probes, write_outputs=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        h2d = (host_bytes + K * M * 8) / iterations + K * M
        diff = max(np.linalg.norm(x1 - x2) / max(np.linalg.norm(x1), 1e-300) for x1, x2 in zip(xs, xs2))
        e2e = {"value": iterations / dt, "unit": "it/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(M * 8 + 512),
               "what": "VAMP.load_ld(host LD, fp32 %s) + VAMP.infer(host r) for %d iterations from it=0, wall clock; LD upload + "
                       "layout conversion + symmetry check + per-iteration probe H2D and xhat D2H inside the timed region" % (
                           "ndarray" if cfg["kind"] == "dense" else "scipy CSR", iterations),
               "seconds": dt, "ld_upload_seconds": t_up, "its_per_s_excluding_upload": iterations / max(dt - t_up, 1e-9),
               "max_rel_diff_vs_resident": float(diff), "layout": v2.handle.ld_info(0)["layout"]}
        v2.close()
        del hostR
    elif not a.no_e2e:
        e2e = {"value": None, "unit": "it/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
               "what": "not run: the host copy of the LD matrices (%.0f GB) is not staged for this shape" % (host_bytes / 1e9)}

    # ---- CPU leg + parity on a bounded sample of the same generator
    cpu, parity = None, None
    if not a.no_cpu_baseline:
        cpu, ref, (Rh, rh, Nss, Ms, pr_s, its_c) = cpu_leg(torch, cfg, bench, seed, dev, ncores)
        vs = new_solver(Ms, Nss)
        xg = vs.infer(Rh if K > 1 else Rh[0], list(rh) if K > 1 else rh[0], its_c, cg_maxit=cfg["cg_maxit"], em_prior_maxit=100,
                      learn_gamw=True, lmmse_damp=False, prior_update="em", update_prior_from=1, probes=pr_s, s=0.0)
        xerr = max(float(np.linalg.norm(xg[it].ravel() - ref["xhat1"][it]) / np.linalg.norm(ref["xhat1"][it])) for it in range(its_c))
        serr, cg_equal = 0.0, True
        for it in range(its_c):
            for k in range(K):
                a_ = np.array(vs.history["rows"][it][k][1:7], dtype=np.float64)
                b_ = np.array(ref["rows"][it][k][1:7], dtype=np.float64)
                serr = max(serr, float(np.max(np.abs(a_ - b_) / np.abs(b_))))
                cg_equal = cg_equal and tuple(vs.history["cg_iters"][it][k]) == tuple(ref["cg_iters"][it][k])
        parity = {"xhat_rel_l2_max": xerr, "scalar_rel_max": serr, "cg_iters_equal": bool(cg_equal), "iterations": its_c,
                  "sample_M": Ms, "layout": vs.handle.ld_info(0)["layout"],
                  "what": "GPU (C ABI, host LD upload) vs CPU oracle on the cpu_baseline sample: same LD, r, probes"}
        vs.close()
    sampler.close()
    kernel = {"dense": "k_gather_rows + k_spmm_panel (2-RHS column sweep over the rank's column panel) + k_panel_finish" if rows_part else
                       "k_spmm_psym (2-RHS upper-triangle symmetric dense SpMM) + k_psym_finish",
              "blockdiag": "k_spmm_psym (2-RHS upper-triangle SpMM over the LD blocks' panels) + k_psym_finish"}[cfg["kind"]]
    line = {
        "metric": "VAMP iterations/s", "value": value, "unit": "it/s", "n_gpus": world if sharded else 1, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "%s: %s, cg_maxit=%d, rho=%.1f" % (a.config, cfg["text"], cfg["cg_maxit"], cfg.get("rho", 0.5)), "M": M, "K": K, "layout": layout,
                   "nnz_stored": nnz_all, "nblocks_rank0": infos[0]["nblocks"],
                   "partition": ("%d GPUs, every cohort's dense LD partitioned by rows (column panel R[:, lo:hi] per GPU); the vector "
                                 "pair of all ranks is gathered from peer memory before each product (CG direction update fused "
                                 "into the gather), reductions exchanged in-kernel" % world) if rows_part else
                                ("%d GPUs, whole LD blocks per GPU balanced by sum m_b^2 (shard.partition_blocks); SpMM local, "
                                 "scalar reductions exchanged in-kernel through peer memory" % world) if sharded else "single GPU",
                   "l2_policy": "inputs (%.1f GB of LD in HBM) larger than L2" % (sum(i["nnz_stored"] for i in infos) * 4 / 1e9),
                   "timed_iterations": "VAMP iterations %d..%d of one trajectory" % (a.warmup, iterations - 1),
                   "cg_iters_timed": [[list(hist["cg_iters"][i][k]) for k in range(K)] for i in range(a.warmup, iterations)],
                   "spmm_passes_timed": passes, "alignment_with_truth": aligns[-1], "gen_seconds": t_gen,
                   "note": "single-GPU shape; under torchrun only rank 0 runs it" if (world > 1 and not sharded) else None},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "kernel": kernel, "bytes_per_launch": bytes_pass, "avg_launch_ms": avg_ms, "launches_timed": spmm_launches,
                     "isolated_launch_ms": iso_ms, "isolated_gbs": (bytes_pass / (iso_ms * 1e-3) / 1e9) if iso_ms else None,
                     "per": "GPU (slowest rank)",
                     "bytes_definition": "the rank's fp32 column panel once (4 B per entry) + 16 B per gathered marker + 32 B per own marker" if rows_part
                                         else "upper triangle of the symmetric fp32 store (2 B per matrix entry) + 32 B per marker",
                     "spmm_share_of_step": spmm_ms / ms_total},
        "cpu_baseline": cpu, "parity": parity, "e2e": e2e, "gpu_launches": nlaunch, "clocks": clocks,
    }
    print(json.dumps(line))
    if sharded:
        dist.barrier()
        dist.destroy_process_group()
