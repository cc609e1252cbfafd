"""The kernel instantiation and the shape that bench.py measures, pinned directly to scipy / the oracle.

bench.py's M = 1M run uses the BIG tile of the symmetric half-band kernel (chosen once the stored rows reach
sm_count * 512, about 75.8k on a B200) and, in the solver loop, the fused conjugate-gradient step built on it.
The goldens stop at M = 40k (small tile), so these tests compare exactly those kernels
  * with scipy's own matrix-vector product at M >= 80k rows for w in {250, 500, 1000}, on one rank and on 2 / 4
    row-partitioned ranks (every rank's shard above the tile switch), and
  * with the CPU oracle (scipy CG semantics) on a 3-iteration VAMP trajectory at M = 80k, w = 500 built by the
    same generator as the benchmark workload, CG iteration counts asserted.
"""
import numpy as np
import pytest
import scipy.sparse

from golden_util import rel_err, rel_l2
from test_gpu_sharded import run_ranks

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nat():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import sgv_native
    return sgv_native


def _upper_band(M, w, seed):
    """fp32 upper diagonals U[d, i] = R[i, i+d] (zero where i+d >= M) of a random symmetric banded matrix."""
    rng = np.random.default_rng(seed)
    U = rng.standard_normal((w + 1, M), dtype=np.float32)
    for d in range(1, w + 1):
        U[d, M - d:] = 0.0
    return U


def _scipy_sym_band(U):
    """The symmetric matrix with upper diagonals U as a scipy DIA matrix (fp64): scipy's matvec is the reference."""
    w, M = U.shape[0] - 1, U.shape[1]
    data = np.zeros((2 * w + 1, M))
    for d in range(0, w + 1):
        data[w + d, d:] = U[d, : M - d]          # offset +d: data[k, j] = R[j-d, j] = U[d, j-d]
        if d:
            data[w - d, : M - d] = U[d, : M - d]  # offset -d: data[k, j] = R[j+d, j] = U[d, j]
    return scipy.sparse.dia_matrix((data, np.arange(-w, w + 1)), shape=(M, M))


def _dsym_device(torch, U, lo, hi, ext, dev):
    """Rows [lo-ext, hi) of the half band in the tiled layout of sgv_ld_adopt_dsym (diagonal halved; extension rows keep
    only their couplings to the rank's own rows)."""
    import ldgen
    w = U.shape[0] - 1
    glo, n = lo - ext, hi - (lo - ext)
    Dp = (w + 1 + 3) // 4 * 4
    ldb = (n + 127) // 128 * 128
    T = torch.zeros((Dp, ldb), device=dev, dtype=torch.float32)
    T[: w + 1, :n] = torch.from_numpy(np.ascontiguousarray(U[:, glo:hi])).to(dev)
    T[0] *= 0.5
    if ext:
        jj = torch.arange(ext, device=dev)[None, :]
        dd = torch.arange(Dp, device=dev)[:, None]
        T[:, :ext] *= (jj + dd >= ext).to(torch.float32)
    return ldgen.dsym_tile(torch, T), ldb


@pytest.mark.parametrize("w", [250, 500, 1000])
def test_bigtile_spmm_vs_scipy(nat, w):
    import torch
    M = 80_000                                     # 80k rows: above the tile switch (sm_count * 512 = 75 776 on B200)
    U = _upper_band(M, w, seed=w)
    R = _scipy_sym_band(U)
    dev = torch.device("cuda", 0)
    T, ldb = _dsym_device(torch, U, 0, M, 0, dev)
    h = nat.Handle()
    h.configure(M, 1)
    h.adopt_dsym(0, T.data_ptr(), w, ldb, 0)
    assert h.ld_info(0)["layout"] == "dsym"
    X = np.random.default_rng(1).standard_normal((M, 2))
    Y = h.spmm(0, X, alpha=1.7, beta=-0.3)
    assert rel_l2(Y, 1.7 * (R @ X) - 0.3 * X) < 1e-13
    assert rel_l2(h.spmm(0, X[:, 1].copy()), R @ X[:, 1]) < 1e-13
    # the same matrix through the host upload path (scipy DIA arrays -> device conversion into the tiled half band)
    Rf = R.copy()
    Rf.data = Rf.data.astype(np.float32)
    h._ck(h.upload_dia(0, Rf.data, Rf.offsets, s=0.0, layout=nat.LAYOUT_AUTO))
    assert h.ld_info(0)["layout"] == "dsym" and h.ld_info(0)["bandwidth"] == w
    assert np.array_equal(h.spmm(0, X, alpha=1.7, beta=-0.3), Y)
    h.close()


@pytest.mark.parametrize("world", [2, 4])
def test_bigtile_spmm_sharded_vs_scipy(nat, world):
    import torch
    import shard as shd
    w = 500
    M = world * 76_800                             # every shard (plus its extension rows) above the tile switch
    U = _upper_band(M, w, seed=world)
    R = _scipy_sym_band(U)
    X = np.random.default_rng(2).standard_normal((M, 2))
    bounds = shd.partition_rows(M, world)

    def fn(sh, dev):
        lo, hi = bounds[sh.rank]
        h = nat.Handle(device=dev)
        h.configure_part(M, 1, sh.rank, world, lo, hi, True)
        shd.attach_peers(h, sh)
        ext = h.dsym_extension(w)
        assert ext == (0 if sh.rank == 0 else (w + 255) // 256 * 256)
        T, ldb = _dsym_device(torch, U, lo, hi, ext, torch.device("cuda", dev))
        h.adopt_dsym(0, T.data_ptr(), w, ldb, ext)
        h.spmm_stage(X[lo:hi])
        sh.barrier()
        Y = h.spmm_run(0, 2, alpha=1.3, beta=-0.7)
        sh.barrier()
        h.close()
        return Y

    Y = np.concatenate(run_ranks(world, fn), axis=0)
    assert rel_l2(Y, 1.3 * (R @ X) - 0.7 * X) < 1e-13


def test_bench_shape_trajectory_vs_oracle(nat):
    """M = 80k, w = 500 from the benchmark's own generator and parameters: the big-tile fused-CG path against the CPU
    oracle, 3 VAMP iterations, same probes; the CG iteration counts must be equal."""
    import torch
    import bench
    import sgvamp
    from oracle import sgvamp_oracle as orc
    M, w, its, seed = 80_000, 500, 3, 5
    dev = torch.device("cuda", 0)
    U, ldb, band, r, x0, _ = bench.build_problem(torch, M, w, seed, dev)
    p = bench.vamp_params(M)
    probes = bench.make_probes(its, M, seed)
    v = sgvamp.VAMP(N=bench.n_gwas(M), Nt=bench.n_gwas(M), M=M, K=1, rho=p["rho"], gamw=p["gamw"], gam1=p["gam1"],
                    a=np.array([1.0]), prior_vars=p["prior_vars"], prior_probs=p["prior_probs"], out_dir=None, out_name="t")
    xs = v.infer(sgvamp.DeviceDSYM(U.data_ptr(), w, ldb, 0, keepalive=U), r, its, cg_maxit=p["cg_maxit"],
                 em_prior_maxit=p["em_prior_maxit"], learn_gamw=True, lmmse_damp=False, prior_update="em",
                 update_prior_from=1, probes=probes)
    assert v.handle.ld_info(0)["layout"] == "dsym"
    rows = [v.history["rows"][i][0] for i in range(its)]
    cg = [tuple(v.history["cg_iters"][i][0]) for i in range(its)]
    v.close()
    R = bench.band_to_scipy_dia(band.cpu().numpy(), M, w)          # Rused exactly as stored (fp32-representable)
    del band
    o = orc.VAMPOracle([bench.n_gwas(M)], M, p["rho"], p["gamw"], p["gam1"], p["prior_vars"], p["prior_probs"])
    ref = o.infer([R], [r], its, cg_maxit=p["cg_maxit"], em_prior_maxit=p["em_prior_maxit"], learn_gamw=True,
                  lmmse_damp=False, prior_update="em", update_prior_from=1, probe_fn=lambda k, it, M_: probes[k, it])
    for it in range(its):
        assert cg[it] == tuple(ref["cg_iters"][it][0]), (it, cg[it], ref["cg_iters"][it][0])
        assert rel_l2(xs[it], ref["xhat1"][it]) <= 1e-4
        assert rel_err(rows[it][1:6], ref["rows"][it][0][1:6]) <= 1e-4
        # LD values are fp32-representable, so only summation order differs: far inside the 1e-4 bar
        assert rel_l2(xs[it], ref["xhat1"][it]) <= 1e-8, (it, rel_l2(xs[it], ref["xhat1"][it]))
    assert cg[0][0] > 10                                           # a real solve, not a trivially converged one
