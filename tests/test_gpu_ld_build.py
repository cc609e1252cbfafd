"""On-GPU LD construction (sgv_ld_build_banded) against the reference's own recipe, simulation/sim_gen_phen_mult.py:39-55:
X ~ genotypes {0,1,2}, X = (X - mean) / std, X /= sqrt(N), r = X^T y, R = X^T X - evaluated with numpy in fp64 on the same
genotypes and restricted to the band (|i-j| <= w, optional Bartlett taper, Rused = (1-s) R + s I as src/main.py:265).
LD values are stored fp32, so products agree to fp32 rounding of the matrix (~1e-7); r is fp64 throughout."""
import numpy as np
import pytest
import scipy.sparse

from golden_util import rel_l2
from test_gpu_sharded import run_ranks

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nat():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import sgv_native
    return sgv_native


def _recipe(G, y, w, s, taper):
    """numpy restatement of the recipe on genotypes G (N x M): banded Rused (scipy CSR, fp64) and r."""
    N, M = G.shape
    X = G.astype(np.float64)
    X = (X - X.mean(axis=0)) / X.std(axis=0)
    X /= np.sqrt(N)
    r = X.T @ y
    R = X.T @ X
    ii, jj = np.meshgrid(np.arange(M), np.arange(M), indexing="ij")
    d = np.abs(ii - jj)
    R = np.where(d <= w, R, 0.0)
    if taper:
        R = R * np.where(d <= w, 1.0 - d / (w + 1.0), 0.0)
    np.fill_diagonal(R, 1.0)
    R = (1 - s) * R + s * np.eye(M)
    return scipy.sparse.csr_matrix(R), r


def _genotypes(N, M, seed):
    rng = np.random.default_rng(seed)
    maf = rng.uniform(0.05, 0.5, M)
    base = rng.random((N, 1)) < 0.5                       # some shared structure so that neighbours correlate
    G = (rng.random((N, M)) < maf[None, :]).astype(np.int8) + ((rng.random((N, M)) < maf[None, :]) & base).astype(np.int8)
    return G


def _marker_major(G):
    N, M = G.shape
    ldg = (N + 15) // 16 * 16
    Gt = np.zeros((M, ldg), dtype=np.int8)
    Gt[:, :N] = G.T
    return Gt


@pytest.mark.parametrize("N,M,w,s,taper", [(300, 700, 40, 0.0, False), (1000, 2000, 257, 0.1, True), (77, 130, 5, 0.1, True),
                                           (513, 1500, 129, 0.0, True)])
def test_ld_build_matches_recipe(nat, N, M, w, s, taper):
    G = _genotypes(N, M, seed=N + M)
    y = np.random.default_rng(1).standard_normal(N)
    R, r = _recipe(G, y, w, s, taper)
    h = nat.Handle()
    h.configure(M, 1)
    r_gpu = h.build_banded(0, _marker_major(G), N, w, s=s, taper=taper, y=y)
    info = h.ld_info(0)
    assert info["layout"] == "dsym" and info["bandwidth"] == w
    assert rel_l2(r_gpu, r) < 1e-12
    X = np.random.default_rng(2).standard_normal((M, 2))
    assert rel_l2(h.spmm(0, X, alpha=1.3, beta=-0.2), 1.3 * (R @ X) - 0.2 * X) < 5e-7      # fp32 storage of R
    # the same matrix uploaded through the host path gives the same stored values (to the fp32 rounding of either route)
    Y1 = h.spmm(0, X)
    h._ck(h.upload_csr(0, R.indptr, R.indices, R.data, s=0.0, layout=nat.LAYOUT_DSYM))
    assert rel_l2(h.spmm(0, X), Y1) < 3e-7
    h.close()


def test_ld_build_device_pointer_and_monomorphic_marker(nat):
    import torch
    N, M, w = 400, 900, 64
    G = _genotypes(N, M, seed=3)
    G[:, 17] = 1                                          # monomorphic: zero variance -> left uncoupled, r = 0
    y = np.random.default_rng(4).standard_normal(N)
    Gt = torch.from_numpy(_marker_major(G)).cuda()
    h = nat.Handle()
    h.configure(M, 1)
    r_gpu = h.build_banded(0, None, N, w, s=0.0, taper=True, y=y, device_ptr=Gt.data_ptr(), nmark=M, ldg=Gt.shape[1])
    assert r_gpu[17] == 0.0 and np.all(np.isfinite(r_gpu))
    e = np.zeros((M, 1))
    e[17] = 1.0
    col = h.spmm(0, e[:, 0].copy())
    assert abs(col[17] - 1.0) < 1e-7 and np.abs(np.delete(col, 17)).max() == 0.0
    h.close()


@pytest.mark.parametrize("world", [2, 3])
def test_ld_build_row_partition(nat, world):
    """Every rank builds its own rows (plus the extension rows before them) from the genotype window it needs."""
    import shard as shd
    N, M, w, s = 500, 3000, 100, 0.1
    G = _genotypes(N, M, seed=9)
    y = np.random.default_rng(5).standard_normal(N)
    R, r = _recipe(G, y, w, s, True)
    Gt = _marker_major(G)
    X = np.random.default_rng(6).standard_normal((M, 2))
    bounds = shd.partition_rows(M, world)

    def fn(sh, dev):
        lo, hi = bounds[sh.rank]
        h = nat.Handle(device=dev)
        h.configure_part(M, 1, sh.rank, world, lo, hi, True)
        shd.attach_peers(h, sh)
        E = h.dsym_extension(w)
        g0, g1 = max(0, lo - E), min(M, hi + w)
        rl = h.build_banded(0, np.ascontiguousarray(Gt[g0:g1]), N, w, s=s, taper=True, y=y, g0=g0)
        h.spmm_stage(X[lo:hi])
        sh.barrier()
        Y = h.spmm_run(0, 2)
        sh.barrier()
        h.close()
        return Y, rl

    out = run_ranks(world, fn)
    Y = np.concatenate([o[0] for o in out], axis=0)
    rr = np.concatenate([o[1] for o in out])
    assert rel_l2(Y, R @ X) < 5e-7 and rel_l2(rr, r) < 1e-12


def test_ld_build_tensor_core_and_dp4a_kernels_identical(nat, monkeypatch):
    """The Gram sums are exact integers in both kernels (tcgen05.mma kind::i8 with TMEM accumulators; IDP4A on the CUDA
    cores), and the epilogue is the same arithmetic: the two constructed matrices must be bit-identical."""
    N, M, w = 1000, 3000, 300
    G = _genotypes(N, M, seed=21)
    Gt = _marker_major(G)
    X = np.random.default_rng(3).standard_normal((M, 2))
    outs = []
    for flag in ("0", "1"):
        monkeypatch.setenv("SGV_LD_DP4A", flag)
        h = nat.Handle()
        h.configure(M, 1)
        h.build_banded(0, Gt, N, w, s=0.1, taper=True)
        outs.append(h.spmm(0, X))
        h.close()
    monkeypatch.delenv("SGV_LD_DP4A")
    assert np.array_equal(outs[0], outs[1])
