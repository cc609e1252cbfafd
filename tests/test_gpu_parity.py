"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and against the
golden vectors produced by the unmodified reference.

Bar (BASELINE.json north_star): per-iteration xhat rel-L2 <= 1e-4; gamw/gam1/gam2/alpha1/alpha2
within 1e-4 relative, same injected probes.  Kernel-level checks use much tighter bounds.
"""
import os
import tempfile

import numpy as np
import pytest
import scipy.sparse

from golden_util import ALL_CASES, UNSTABLE, load_case, rel_err, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nat():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import sgv_native
    return sgv_native


@pytest.fixture(scope="module")
def orc():
    from oracle import sgvamp_oracle
    return sgvamp_oracle


def _rand_sym_band(M, w, seed, fill=1.0):
    rng = np.random.default_rng(seed)
    offs = np.arange(0, w + 1)
    diags = []
    for o in offs:
        d = rng.standard_normal(M - o).astype(np.float32).astype(np.float64)
        if fill < 1.0 and o > 0:
            d *= (rng.random(M - o) < fill)
        diags.append(d)
    U = scipy.sparse.diags(diags, offs, shape=(M, M), format="csr")
    R = (U + scipy.sparse.triu(U, 1).T).tocsr()
    R.eliminate_zeros()
    R.sort_indices()
    return R


# ---------------------------------------------------------------------------------------------
# SpMM, every layout
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,w", [(1, 0), (7, 3), (129, 5), (1000, 40), (5000, 257), (40000, 33)])
@pytest.mark.parametrize("layout", ["dia", "dsym", "csr"])
def test_spmm_banded(nat, M, w, layout):
    w = min(w, M - 1)
    R = _rand_sym_band(M, w, seed=M + w)
    h = nat.Handle()
    h.configure(M, 1)
    h._ck(h.upload_csr(0, R.indptr, R.indices, R.data, s=0.0,
                       layout={"dia": nat.LAYOUT_DIA, "dsym": nat.LAYOUT_DSYM, "csr": nat.LAYOUT_CSR}[layout]))
    info = h.ld_info(0)
    assert info["layout"] == layout
    rng = np.random.default_rng(1)
    X = rng.standard_normal((M, 2))
    Y = h.spmm(0, X, alpha=1.7, beta=-0.3)
    ref = 1.7 * (R @ X) - 0.3 * X
    assert rel_l2(Y, ref) < 1e-13
    y1 = h.spmm(0, X[:, 0].copy())
    assert rel_l2(y1, R @ X[:, 0]) < 1e-13
    h.close()


@pytest.mark.parametrize("M", [1, 5, 130, 515, 2049])
def test_spmm_dense(nat, M):
    rng = np.random.default_rng(M)
    B = rng.standard_normal((M, M))
    R = (B + B.T).astype(np.float32).astype(np.float64)
    h = nat.Handle()
    h.configure(M, 1)
    h.upload_dense(0, R, s=0.0)
    assert h.ld_info(0)["layout"] == "dense"
    X = rng.standard_normal((M, 2))
    Y = h.spmm(0, X, alpha=0.9, beta=2.0)
    assert rel_l2(Y, 0.9 * (R @ X) + 2.0 * X) < 1e-13
    # fp32 input path
    h.upload_dense(0, R.astype(np.float32), s=0.0)
    assert rel_l2(h.spmm(0, X), R @ X) < 1e-13
    h.close()


def test_spmm_blockdiag_and_auto_detection(nat):
    rng = np.random.default_rng(3)
    sizes = [1, 37, 512, 513, 260, 4, 1030]
    blocks = []
    for b in sizes:
        B = rng.standard_normal((b, b))
        blocks.append((B + B.T + np.eye(b)).astype(np.float32).astype(np.float64))
    R = scipy.sparse.block_diag(blocks, format="csr")
    R.sort_indices()
    M = R.shape[0]
    h = nat.Handle()
    h.configure(M, 1)
    h._ck(h.upload_csr(0, R.indptr, R.indices, R.data))
    info = h.ld_info(0)
    assert info["layout"] == "blockdiag" and info["nblocks"] == len(sizes)
    X = rng.standard_normal((M, 2))
    assert rel_l2(h.spmm(0, X), R @ X) < 1e-13
    # banded input is detected as DIA, a full matrix in CSR clothing as dense, thin random as CSR
    Rb = _rand_sym_band(3000, 20, 5)
    h.configure(3000, 1)
    h._ck(h.upload_csr(0, Rb.indptr, Rb.indices, Rb.data))
    assert h.ld_info(0)["layout"] == "dsym" and h.ld_info(0)["bandwidth"] == 20
    # a banded matrix that is not symmetric cannot use the half band: full band instead, explicit dsym refuses
    Ra = Rb.copy().tolil()
    Ra[10, 12] = 0.25 if Ra[12, 10] != 0.25 else 0.75      # fp32-representable, differs from its mirror
    Ra = Ra.tocsr()
    Ra.sort_indices()
    h._ck(h.upload_csr(0, Ra.indptr, Ra.indices, Ra.data))
    assert h.ld_info(0)["layout"] == "dia"
    Xa = rng.standard_normal((3000, 2))
    assert rel_l2(h.spmm(0, Xa), Ra @ Xa) < 1e-13
    assert h.upload_csr(0, Ra.indptr, Ra.indices, Ra.data, layout=nat.LAYOUT_DSYM) != 0
    Rs = _rand_sym_band(3000, 400, 6, fill=0.02)
    h._ck(h.upload_csr(0, Rs.indptr, Rs.indices, Rs.data))
    assert h.ld_info(0)["layout"] == "csr"
    X = rng.standard_normal((3000, 2))
    assert rel_l2(h.spmm(0, X), Rs @ X) < 1e-13
    D = scipy.sparse.csr_matrix(blocks[2])
    h.configure(512, 1)
    h._ck(h.upload_csr(0, D.indptr, D.indices, D.data))
    assert h.ld_info(0)["layout"] == "dense"
    h.close()


@pytest.mark.parametrize("layout", ["auto", "dia", "dsym"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_upload_scipy_dia_format(nat, layout, dtype):
    """LD handed over in scipy's DIA format (no index arrays; only the upper diagonals travel for the symmetric
    half band, the lower ones are verified on the device): same products as the CSR route, Rused applied."""
    M, w, s = 3000, 37, 0.1
    R = _rand_sym_band(M, w, seed=8)
    Rd = R.todia()
    Rd.data = Rd.data.astype(dtype)
    lay = {"auto": nat.LAYOUT_AUTO, "dia": nat.LAYOUT_DIA, "dsym": nat.LAYOUT_DSYM}[layout]
    h = nat.Handle()
    h.configure(M, 1)
    X = np.random.default_rng(1).standard_normal((M, 2))
    h._ck(h.upload_dia(0, Rd.data, Rd.offsets, s=0.0, layout=lay))
    info = h.ld_info(0)
    assert info["layout"] == ("dia" if layout == "dia" else "dsym") and info["bandwidth"] == w
    assert rel_l2(h.spmm(0, X), R @ X) < 1e-13                         # values are fp32-representable: exact storage
    h._ck(h.upload_dia(0, Rd.data, Rd.offsets, s=s, layout=lay))       # Rused rounds once to fp32
    ref = (1 - s) * (R @ X) + s * X
    assert rel_l2(h.spmm(0, X), ref) < 1e-6
    # a column window of the DIA arrays (what one rank of a row partition passes) gives the same matrix
    if layout == "dia":                                                 # (columns outside the window read as zero)
        h._ck(h.upload_dia(0, np.ascontiguousarray(Rd.data[:, 5:M - 3]), Rd.offsets, s=0.0, layout=lay, col0=5))
        Rw = R.tolil()
        Rw[:, :5] = 0
        Rw[:, M - 3:] = 0
        assert rel_l2(h.spmm(0, X), Rw.tocsr() @ X) < 1e-13
    # not symmetric: auto falls back to the full band, explicit dsym refuses, assume_symmetric uses the upper half
    Ra = R.tolil()
    Ra[100, 103] = 0.25 if Ra[103, 100] != 0.25 else 0.75
    Ra = Ra.tocsr().todia()
    if layout == "auto":
        h._ck(h.upload_dia(0, Ra.data, Ra.offsets, s=0.0, layout=lay))
        assert h.ld_info(0)["layout"] == "dia"
        assert rel_l2(h.spmm(0, X), Ra @ X) < 1e-13
    if layout == "dsym":
        assert h.upload_dia(0, Ra.data, Ra.offsets, s=0.0, layout=lay) != 0
        h._ck(h.upload_dia(0, Ra.data, Ra.offsets, s=0.0, layout=lay, assume_symmetric=True))
        Rsym = scipy.sparse.triu(Ra.tocsr(), 0)
        Rsym = (Rsym + scipy.sparse.triu(Ra.tocsr(), 1).T).tocsr()
        assert rel_l2(h.spmm(0, X), Rsym @ X) < 1e-13
    h.close()


def test_symmetric_and_full_panel_kernels_agree(nat, monkeypatch):
    """Dense / block-diagonal LD goes through the upper-triangle kernel (spmm_psym.cu) by default and through
    the full-panel kernel with SGV_PANEL_FULL=1: same products, to rounding."""
    rng = np.random.default_rng(4)
    for sizes in ([1], [5], [515], [2049], [37, 512, 513, 260, 4, 1030, 1]):
        blocks = []
        for b in sizes:
            B = rng.standard_normal((b, b))
            blocks.append((B + B.T).astype(np.float32).astype(np.float64))
        R = scipy.sparse.block_diag([scipy.sparse.csr_matrix(b) for b in blocks], format="csr")
        R.sort_indices()
        M = R.shape[0]
        h = nat.Handle()
        h.configure(M, 1)
        if len(sizes) == 1:
            h.upload_dense(0, blocks[0], s=0.0)
        else:
            h._ck(h.upload_csr(0, R.indptr, R.indices, R.data, layout=nat.LAYOUT_BLOCKDIAG))
        X = rng.standard_normal((M, 2))
        monkeypatch.delenv("SGV_PANEL_FULL", raising=False)
        sym = h.spmm(0, X, alpha=1.1, beta=0.3)
        monkeypatch.setenv("SGV_PANEL_FULL", "1")
        full = h.spmm(0, X, alpha=1.1, beta=0.3)
        monkeypatch.delenv("SGV_PANEL_FULL", raising=False)
        ref = 1.1 * (R @ X) + 0.3 * X
        assert rel_l2(full, ref) < 1e-13 and rel_l2(sym, ref) < 1e-13, sizes
        h.close()


def test_nonsymmetric_ld_is_not_mirrored_silently(nat):
    """The reference computes R @ v for whatever R it is given (src/sgvamp.py:316,352).  The half-band and
    upper-triangle kernels use symmetry, so every upload path verifies it on the device and routes a non-symmetric R
    to a kernel that keeps R as given: dense / block-diagonal -> transposed store + full-panel kernel, banded -> full
    band; an explicit request for the half band refuses."""
    rng = np.random.default_rng(12)
    # dense, not symmetric
    M = 515
    R = rng.standard_normal((M, M)).astype(np.float32).astype(np.float64)
    X = rng.standard_normal((M, 2))
    h = nat.Handle()
    h.configure(M, 1)
    h.upload_dense(0, R, s=0.0)
    assert rel_l2(h.spmm(0, X, alpha=1.2, beta=0.4), 1.2 * (R @ X) + 0.4 * X) < 1e-13
    # one entry off in an otherwise symmetric dense matrix
    Rs = (R + R.T).astype(np.float32).astype(np.float64)
    Rs[7, 300] = 0.5 if Rs[300, 7] != 0.5 else 0.75
    h.upload_dense(0, Rs, s=0.0)
    assert rel_l2(h.spmm(0, X), Rs @ X) < 1e-13
    # block-diagonal from CSR with one non-symmetric block
    blocks = []
    for b in (40, 260, 33):
        B = rng.standard_normal((b, b))
        blocks.append((B + B.T).astype(np.float32).astype(np.float64))
    blocks[1][3, 200] = 0.25
    Rb = scipy.sparse.block_diag([scipy.sparse.csr_matrix(b) for b in blocks], format="csr")
    Rb.sort_indices()
    Mb = Rb.shape[0]
    h.configure(Mb, 1)
    h._ck(h.upload_csr(0, Rb.indptr, Rb.indices, Rb.data, layout=nat.LAYOUT_BLOCKDIAG))
    Xb = rng.standard_normal((Mb, 2))
    assert rel_l2(h.spmm(0, Xb), Rb @ Xb) < 1e-13
    # banded CSR holding only the upper triangle: not symmetric -> full band; explicit half band refuses
    Mt, w = 3000, 20
    Rt = scipy.sparse.triu(_rand_sym_band(Mt, w, 31), 0).tocsr()
    Rt.sort_indices()
    h.configure(Mt, 1)
    h._ck(h.upload_csr(0, Rt.indptr, Rt.indices, Rt.data))
    assert h.ld_info(0)["layout"] == "dia"
    Xt = rng.standard_normal((Mt, 2))
    assert rel_l2(h.spmm(0, Xt), Rt @ Xt) < 1e-13
    assert h.upload_csr(0, Rt.indptr, Rt.indices, Rt.data, layout=nat.LAYOUT_DSYM) != 0
    # the same in scipy's DIA container (only non-negative offsets stored)
    Rd = Rt.todia()
    assert Rd.offsets.min() >= 0
    h._ck(h.upload_dia(0, Rd.data, Rd.offsets))
    assert h.ld_info(0)["layout"] == "dia"
    assert rel_l2(h.spmm(0, Xt), Rt @ Xt) < 1e-13
    assert h.upload_dia(0, Rd.data, Rd.offsets, layout=nat.LAYOUT_DSYM) != 0
    h._ck(h.upload_dia(0, Rd.data, Rd.offsets, layout=nat.LAYOUT_DSYM, assume_symmetric=True))   # declared symmetric: mirrored
    Rm = (Rt + scipy.sparse.triu(Rt, 1).T).tocsr()
    assert rel_l2(h.spmm(0, Xt), Rm @ Xt) < 1e-13
    h.close()


def test_regularisation_at_upload(nat):
    """Rused = (1-s) R + s I (src/main.py:265) on every layout, including absent diagonal entries."""
    rng = np.random.default_rng(9)
    M, s = 700, 0.1
    R = _rand_sym_band(M, 12, 11)
    R.setdiag(1.0)
    R.sort_indices()
    X = rng.standard_normal((M, 2))
    ref = ((1 - s) * R + s * scipy.sparse.identity(M)) @ X
    h = nat.Handle()
    h.configure(M, 1)
    for lay in (nat.LAYOUT_DIA, nat.LAYOUT_CSR, nat.LAYOUT_DENSE, nat.LAYOUT_BLOCKDIAG):
        h._ck(h.upload_csr(0, R.indptr, R.indices, R.data, s=s, layout=lay))
        assert rel_l2(h.spmm(0, X), ref) < 1e-7, lay     # fp32 rounding of (1-s)*R+s
    h.upload_dense(0, R.toarray(), s=s)
    assert rel_l2(h.spmm(0, X), ref) < 1e-7
    # missing diagonal: DIA/dense fill s; CSR reports -3
    R0 = R.copy()
    R0.setdiag(0.0)
    R0.eliminate_zeros()
    ref0 = ((1 - s) * R0 + s * scipy.sparse.identity(M)) @ X
    h._ck(h.upload_csr(0, R0.indptr, R0.indices, R0.data, s=s, layout=nat.LAYOUT_DIA))
    assert rel_l2(h.spmm(0, X), ref0) < 1e-7
    assert h.upload_csr(0, R0.indptr, R0.indices, R0.data, s=s, layout=nat.LAYOUT_CSR) == -3
    h.close()


# ---------------------------------------------------------------------------------------------
# denoiser / EM / Lagrangian kernels against the oracle
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("K,L,M", [(1, 2, 1000), (1, 4, 4097), (3, 3, 777), (2, 8, 50000)])
def test_denoise_em_lagrangian(nat, orc, K, L, M):
    rng = np.random.default_rng(K * 100 + L)
    N_list = list(rng.integers(500, 3000, K).astype(float))
    Nt = sum(N_list)
    a = np.array(N_list) / Nt
    pv = [0.0] + list(np.sort(rng.uniform(1e-4, 1e-2, L - 1)))
    pp = [0.95] + [0.05 / (L - 1)] * (L - 1)
    prior = orc.Prior(pv, pp, Nt)
    r1s = rng.standard_normal((K, M)) * 3.0
    r1s[:, ::50] *= 10
    gam1s = rng.uniform(0.05, 2.0, K)
    xh_prev = rng.standard_normal(M)
    h = nat.Handle()
    h.configure(M, K)
    h.set_weights(a)
    h.set_prior(prior.lam, prior.omegas, prior.sigmas)
    for k in range(K):
        h.set_vec(k, nat.VEC_R1, r1s[k])
    xo, dfac = orc.denoise_all(r1s, gam1s, a, prior)
    xl, dl = orc.denoise_all(r1s[:, :64], gam1s, a, prior, per_marker=True)
    assert rel_l2(xo[:64], xl) < 1e-13
    dmean = h.denoise(gam1s, 0.5, False)
    assert rel_l2(h.get_vec(0, nat.VEC_XHAT1), xo) < 1e-12
    assert abs(dmean - dfac.mean()) <= 1e-12 * abs(dfac.mean())
    h.set_vec(0, nat.VEC_XHAT1, xh_prev)
    h.denoise(gam1s, 0.3, True)
    assert rel_l2(h.get_vec(0, nat.VEC_XHAT1), 0.3 * xo + 0.7 * xh_prev) < 1e-12
    # one EM pass, then the EM loop
    p1 = orc.Prior(pv, pp, Nt)
    orc.prior_update_em(r1s, gam1s, a, p1)
    lam, om, steps, rel = h.prior_em(gam1s, 1, 0.0, L - 1)
    assert abs(lam - p1.lam) < 1e-12 * p1.lam and rel_err(om, p1.omegas) < 1e-11
    p2 = orc.Prior(pv, pp, Nt)
    st_o, rel_o = orc.em_loop(r1s, gam1s, a, p2, 100)
    h.set_prior(prior.lam, prior.omegas, prior.sigmas)
    lam, om, steps, rel = h.prior_em(gam1s, 100, 1e-6, L - 1)
    assert steps == st_o
    assert abs(lam - p2.lam) < 1e-10 * p2.lam and rel_err(om, p2.omegas) < 1e-9
    # Lagrangian residual
    omega0 = np.concatenate([[1 - prior.lam], prior.lam * prior.omegas])
    sigma2 = np.concatenate([[1e-16], prior.sigmas])
    x = np.concatenate([omega0 * rng.uniform(0.8, 1.2, L), [0.7]])
    y_o = orc.lagrangian_der(x, omega0, sigma2, r1s, gam1s, a, L)
    y_g = h.lagrangian(gam1s, x, omega0, sigma2)
    assert rel_err(y_g, y_o) < 1e-10
    h.close()


# ---------------------------------------------------------------------------------------------
# CG with scipy semantics
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("maxit", [0, 1, 3, 500])
@pytest.mark.parametrize("warm", [False, True])
def test_lmmse_cg_against_oracle(nat, orc, maxit, warm):
    M = 3000
    rng = np.random.default_rng(17)
    R = _rand_sym_band(M, 25, 21)
    R = (R @ R.T).tocsr() * (1.0 / 50)              # PSD, bandwidth 50
    R.data = R.data.astype(np.float32).astype(np.float64)
    R.sort_indices()
    gamw, gam2, alpha1, rho = 2.5, 0.8, 0.3, 0.5
    xhat1, r1, xty = rng.standard_normal(M), rng.standard_normal(M), rng.standard_normal(M)
    x2p = rng.standard_normal(M) * warm
    sgp = rng.standard_normal(M) * warm
    u = rng.integers(0, 2, M) * 2 - 1
    h = nat.Handle()
    h.configure(M, 1)
    h._ck(h.upload_csr(0, R.indptr, R.indices, R.data))
    h.set_xty(0, xty)
    h.set_vec(0, nat.VEC_XHAT1, xhat1)
    h.set_vec(0, nat.VEC_R1, r1)
    h.set_vec(0, nat.VEC_XHAT2, x2p)
    h.set_vec(0, nat.VEC_SIGMA2U, sgp)
    for damp in (False, True):
        h.set_vec(0, nat.VEC_XHAT2, x2p)
        h.set_vec(0, nat.VEC_SIGMA2U, sgp)
        out = h.lmmse(0, gamw, gam2, alpha1, rho, maxit, damp, True, not warm, u)
        r2 = (xhat1 - alpha1 * r1) / (1 - alpha1)
        mu2 = gamw * xty + gam2 * r2
        mv = lambda v: gamw * (R @ v) + gam2 * v
        x2, i1, n1 = orc.cg(mv, mu2, x2p, maxit)
        sg, i2, n2 = orc.cg(mv, u, sgp, maxit)
        if damp:
            x2 = rho * x2 + (1 - rho) * x2p
        assert (out.cg_iters[0], out.cg_iters[1]) == (n1, n2)
        assert (out.cg_info[0], out.cg_info[1]) == (i1, i2)
        assert rel_l2(h.get_vec(0, nat.VEC_R2), r2) < 1e-14
        assert rel_l2(h.get_vec(0, nat.VEC_XHAT2), x2) < 1e-9
        assert rel_l2(h.get_vec(0, nat.VEC_SIGMA2U), sg) < 1e-9
        assert abs(out.u_sigma2u - u @ sg) < 1e-9 * abs(u @ sg) + 1e-12
        assert abs(out.xhat2_r - x2 @ xty) < 1e-9 * abs(x2 @ xty) + 1e-9
        assert abs(out.xhat2_R_xhat2 - x2 @ (R @ x2)) < 1e-9 * abs(x2 @ (R @ x2)) + 1e-9
        assert abs(out.u_R_sigma2u - u @ (R @ sg)) < 1e-9 * abs(u @ (R @ sg)) + 1e-9
    h.update_r1(0, 0.25)
    assert rel_l2(h.get_vec(0, nat.VEC_R1), (x2 - 0.25 * r2) / 0.75) < 1e-9
    h.close()


def test_cg_zero_rhs(nat):
    """scipy returns (b, 0) when |b| = 0, regardless of x0."""
    M = 500
    R = _rand_sym_band(M, 3, 2)
    R = (R @ R.T).tocsr()
    h = nat.Handle()
    h.configure(M, 1)
    h._ck(h.upload_csr(0, R.indptr, R.indices, R.data))
    h.set_vec(0, nat.VEC_XHAT2, np.ones(M))
    out = h.lmmse(0, 1.0, 0.0, 0.5, 0.5, 20, False, False, False, np.ones(M, dtype=np.int8))
    # xty = xhat1 = r1 = 0  ->  mu2 = 0 -> xhat2 = b = 0, zero iterations
    assert out.cg_iters[0] == 0 and out.cg_info[0] == 0
    assert not h.get_vec(0, nat.VEC_XHAT2).any()
    h.close()


# ---------------------------------------------------------------------------------------------
# full trajectories against the reference's golden outputs
# ---------------------------------------------------------------------------------------------
def run_gpu(c, out_dir=None, layout="auto"):
    import sgvamp
    K, M = c["K"], c["M"]
    N_list = c["N_list"]
    Nt = sum(N_list)
    v = sgvamp.VAMP(N=N_list if K > 1 else N_list[0], Nt=Nt, M=M, K=K, rho=c["rho"], gamw=c["gamw"], gam1=c["gam1"],
                    a=np.array(N_list) / Nt, prior_vars=c["prior_vars"], prior_probs=c["prior_probs"],
                    out_dir=out_dir, out_name="g", comm=None)
    x0 = c["x0"] * np.sqrt(N_list[0]) if "x0" in c else None
    R = c["R"] if K > 1 else c["R"][0]
    r = list(c["r"]) if K > 1 else c["r"][0]
    probes = c["probes"]
    if "rng_seed" in c:
        # the reference's own draw sequence (src/sgvamp.py:326): seed numpy's legacy global RNG and inject nothing
        np.random.seed(c["rng_seed"])
        probes = None
    xs = v.infer(R, r, c["iterations"], x0=x0, cg_maxit=c["cg_maxit"], em_prior_maxit=c["em_prior_maxit"],
                 learn_gamw=c["learn_gamw"], lmmse_damp=c["lmmse_damp"], prior_update=c["prior_update"],
                 update_prior_from=c["update_prior_from"], s=c["s"], probes=probes, layout=layout)
    info = [v.handle.ld_info(k) for k in range(K)]
    hist = v.history
    fin = (float(v.lam), np.array(v.omegas))
    v.close()
    return xs, hist, info, fin


@pytest.mark.parametrize("name", ALL_CASES)
def test_trajectory_matches_reference(nat, name):
    c = load_case(name)
    with tempfile.TemporaryDirectory() as d:
        # the irregular-sparsity case is pinned to the CSR kernel (auto would pick DIA at this fill)
        xs, hist, info, fin = run_gpu(c, out_dir=d, layout="csr" if c["layout"] == "csr" else "auto")
        expect_layout = {"dense": "dense", "banded": "dsym", "blockdiag": "blockdiag", "csr": "csr"}[c["layout"]]
        assert info[0]["layout"] == expect_layout
        tol = 1e-4
        n_check = c["iterations"] if name not in UNSTABLE else 3    # chaotic regimes: first iterations only
        Nt = sum(c["N_list"])
        for it in range(n_check):
            assert rel_l2(xs[it], c["xhat"][it]) <= tol, (name, it, rel_l2(xs[it], c["xhat"][it]))
            dump = np.fromfile(os.path.join(d, "g_xhat_it_%d.bin" % it))
            assert rel_l2(dump, c["xhat_dump"][it]) <= tol
            for k in range(c["K"]):
                r1d = np.fromfile(os.path.join(d, "g_r1_cohort_%d_it_%d.bin" % (k + 1, it)))
                assert rel_l2(r1d, c["r1_dump"][it, k]) <= tol
                row = hist["rows"][it][k]
                assert rel_err(row[1:6], c["rows"][it, k, 1:6]) <= tol, (name, it, k, row, c["rows"][it, k])
                assert abs(row[6] - c["rows"][it, k, 6]) <= tol * c["rows"][it, k, 6]
                if name not in UNSTABLE:
                    assert tuple(hist["cg_iters"][it][k]) == tuple(c["cg_iters"][it, k]), (name, it, k)
                    assert tuple(hist["cg_info"][it][k]) == tuple(c["cg_info"][it, k])
        if name not in UNSTABLE:
            assert abs(fin[0] - c["final_lam"]) <= tol * c["final_lam"]
            assert rel_err(fin[1], c["final_omegas"]) <= tol
        # file format: header, tab delimiter, CRLF (src/sgvamp.py:39-43)
        raw = open(os.path.join(d, "g_cohort_1.csv"), "rb").read()
        assert raw.startswith(b"it\tgamw\tgam1\tgam2\talpha1\talpha2\tlam\r\n")
        assert raw.count(b"\r\n") == c["iterations"] + 1
        if "metrics" in c:
            m = np.array(hist["metrics"])
            assert rel_err(m[:n_check, 1:], c["metrics"][:n_check, 1:]) <= tol


@pytest.mark.parametrize("name", ["banded_L2_em_s01", "dense_K3_L2_em", "dense_noprior_fixedgamw_damp", "blockdiag_L3_em_s01"])
def test_stepwise_and_fused_loops_agree(nat, name, monkeypatch):
    """VAMP.infer runs the loop either stepwise (scalars on the host, as the reference keeps them; always used for the
    MLE prior update and the rank-per-cohort mode) or fused (scalar chain advanced on the device by the kernels'
    finalisers with the reference's operation order): same trajectories, scalars equal to the last bits."""
    c = load_case(name)
    monkeypatch.setenv("SGV_STEPWISE", "1")
    xs_s, hist_s, _, fin_s = run_gpu(c)
    monkeypatch.delenv("SGV_STEPWISE")
    xs_f, hist_f, _, fin_f = run_gpu(c)
    n_check = c["iterations"] if name not in UNSTABLE else 3
    for it in range(n_check):
        assert rel_l2(xs_f[it], xs_s[it]) <= 1e-12, (name, it)
        for k in range(c["K"]):
            assert rel_err(hist_f["rows"][it][k][1:7], hist_s["rows"][it][k][1:7]) <= 1e-12, (name, it, k)
            assert tuple(hist_f["cg_iters"][it][k]) == tuple(hist_s["cg_iters"][it][k])
            assert tuple(hist_f["cg_info"][it][k]) == tuple(hist_s["cg_info"][it][k])
        assert hist_f["em_steps"][it] == hist_s["em_steps"][it]


@pytest.mark.parametrize("name,stepwise", [("banded_L2_em_s01", False), ("banded_L2_em_defaultrng", False),
                                           ("dense_K3_L2_em", False), ("dense_L2_mle", True)])
def test_checkpoint_and_resume(nat, name, stepwise, tmp_path, monkeypatch):
    """A run checkpointed after half of its iterations and continued by a fresh solver follows the reference's golden
    trajectory to the end (warm starts, damping terms, prior, MLE multiplier and - for un-injected probes - the position
    in numpy's legacy RNG sequence all travel in the checkpoint)."""
    import sgvamp
    c = load_case(name)
    if stepwise:
        monkeypatch.setenv("SGV_STEPWISE", "1")
    K, M, its = c["K"], c["M"], c["iterations"]
    half = its // 2
    Nt = sum(c["N_list"])
    path = str(tmp_path / "ck.npz")

    def solver():
        return sgvamp.VAMP(N=c["N_list"] if K > 1 else c["N_list"][0], Nt=Nt, M=M, K=K, rho=c["rho"], gamw=c["gamw"], gam1=c["gam1"],
                           a=np.array(c["N_list"]) / Nt, prior_vars=c["prior_vars"], prior_probs=c["prior_probs"], out_dir=None,
                           out_name="g")

    def run(v, n_it, **kw):
        probes = c["probes"]
        if "rng_seed" in c:
            probes = None
        return v.infer(c["R"] if K > 1 else c["R"][0], list(c["r"]) if K > 1 else c["r"][0], n_it, cg_maxit=c["cg_maxit"],
                       em_prior_maxit=c["em_prior_maxit"], learn_gamw=c["learn_gamw"], lmmse_damp=c["lmmse_damp"],
                       prior_update=c["prior_update"], update_prior_from=c["update_prior_from"], s=c["s"], probes=probes, **kw)

    if "rng_seed" in c:
        np.random.seed(c["rng_seed"])
    v1 = solver()
    run(v1, half, checkpoint_path=path, checkpoint_every=half)
    v1.close()
    np.random.seed(12345)                      # the sequence position must come from the checkpoint, not from this process
    v2 = solver()
    xs = run(v2, its, resume_from=path)
    hist = v2.history
    v2.close()
    assert all(x is None for x in xs[:half])
    for it in range(half, its):
        assert rel_l2(xs[it], c["xhat"][it]) <= 1e-4, (name, it)
        for k in range(K):
            assert rel_err(hist["rows"][it - half][k][1:6], c["rows"][it, k, 1:6]) <= 1e-4
            assert tuple(hist["cg_iters"][it - half][k]) == tuple(c["cg_iters"][it, k])


@pytest.mark.parametrize("layout", ["csr", "dense", "dia"])
def test_banded_case_other_layouts(nat, layout):
    c = load_case("banded_L2_em_s01")
    xs, hist, info, fin = run_gpu(c, layout=layout)
    assert info[0]["layout"] == layout
    for it in range(c["iterations"]):
        assert rel_l2(xs[it], c["xhat"][it]) <= 1e-4


def test_rank_per_cohort_mode_matches(nat):
    """Reference-style comm (rank = cohort): K threads, one VAMP each, host bcast of r1/gam1."""
    import threading
    import sgvamp
    c = load_case("dense_K3_L2_em")
    K, M = c["K"], c["M"]
    Nt = sum(c["N_list"])
    tls = threading.local()

    class Comm:
        def __init__(self):
            self.bar = threading.Barrier(K)
            self.slot = None

        def Get_rank(self):
            return tls.rank

        def Get_size(self):
            return K

        def bcast(self, obj, root=0):
            if tls.rank == root:
                self.slot = obj
            self.bar.wait()
            out = self.slot
            self.bar.wait()
            return out

    comm = Comm()
    res = [None] * K

    def work(k):
        tls.rank = k
        v = sgvamp.VAMP(N=c["N_list"][k], Nt=Nt, M=M, K=K, rho=c["rho"], gamw=c["gamw"], gam1=c["gam1"],
                        a=np.array(c["N_list"]) / Nt, prior_vars=c["prior_vars"], prior_probs=c["prior_probs"],
                        out_dir=None, out_name="g", comm=comm)
        xs = v.infer(c["R"][k], c["r"][k], c["iterations"], cg_maxit=c["cg_maxit"], lmmse_damp=False,
                     prior_update="em", probes=lambda kk, it, M_: c["probes"][kk, it])
        res[k] = (xs, v.history)
        v.close()

    ths = [threading.Thread(target=work, args=(k,)) for k in range(K)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    for k in range(K):
        assert res[k] is not None
        for it in range(c["iterations"]):
            assert rel_l2(res[k][0][it], c["xhat"][it]) <= 1e-4
            assert rel_err(res[k][1]["rows"][it][k][1:6], c["rows"][it, k, 1:6]) <= 1e-4


# ---------------------------------------------------------------------------------------------
# size-independent properties at larger sizes
# ---------------------------------------------------------------------------------------------
def test_spmm_properties_large(nat):
    M, w = 300000, 200
    rng = np.random.default_rng(5)
    import torch
    band = torch.randn((2 * w + 1, M), device="cuda", dtype=torch.float32)
    # symmetrise: band[w+d, i] = band[w-d, i+d]
    for d in range(1, w + 1):
        band[w + d, : M - d] = band[w - d, d:]
        band[w + d, M - d:] = 0
        band[w - d, :d] = 0
    h = nat.Handle()
    h.configure(M, 1)
    h.adopt_dia(0, band.data_ptr(), w, M)
    x, y = rng.standard_normal(M), rng.standard_normal(M)
    Rx, Ry = h.spmm(0, x), h.spmm(0, y)
    both = h.spmm(0, np.stack([x, y], axis=1))
    assert np.array_equal(both[:, 0], Rx) and np.array_equal(both[:, 1], Ry)      # columns independent, deterministic
    assert abs(y @ Rx - x @ Ry) <= 1e-10 * (np.linalg.norm(y) * np.linalg.norm(Rx))  # symmetry
    lin = h.spmm(0, 2.0 * x - 3.0 * y)
    assert rel_l2(lin, 2.0 * Rx - 3.0 * Ry) < 1e-12                                 # linearity
    # spot-check rows against a direct evaluation
    bh = band[:, 1000:1010].cpu().numpy().astype(np.float64)
    for t in range(10):
        i = 1000 + t
        assert abs(Rx[i] - bh[:, t] @ x[i - w: i + w + 1]) < 1e-9 * np.abs(Rx[i]) + 1e-9
    assert np.array_equal(h.spmm(0, x), Rx)                                         # run-to-run determinism
    # the symmetric half-band layout of the same matrix (upper diagonals, padded to a multiple of 4)
    Dp = (w + 1 + 3) // 4 * 4
    ldb = (M + 127) // 128 * 128
    U = torch.zeros((Dp, ldb), device="cuda", dtype=torch.float32)
    U[: w + 1, :M] = band[w:, :]
    U[0] *= 0.5                                                                     # diagonal stored halved
    import ldgen
    U = ldgen.dsym_tile(torch, U)                                                   # tiled layout of sgv_ld_adopt_dsym
    h.adopt_dsym(0, U.data_ptr(), w, ldb, 0)
    assert h.ld_info(0)["layout"] == "dsym"
    Sx = h.spmm(0, x)
    assert rel_l2(Sx, Rx) < 1e-13                                                   # same product, half the bytes
    both2 = h.spmm(0, np.stack([x, y], axis=1))
    assert np.array_equal(both2[:, 0], Sx)
    assert np.array_equal(h.spmm(0, x), Sx)
    assert h.ld_info(0)["bytes_per_pass"] < 0.52 * (4.0 * (2 * w + 1) * M + 32.0 * M)
    h.close()


def test_full_size_trajectory_properties(nat):
    """BASELINE.json's full size (M = 1M banded, w = 500) is out of the oracle's reach, so the trajectory is
    checked through size-independent properties: the symmetric half-band kernels (fused CG) and the full-band
    kernels (classic CG loop) - two independent implementations - must agree on every iteration to rounding,
    with identical CG iteration counts; a rerun is bit-identical; the estimate aligns with the planted signal."""
    import torch
    import bench
    import sgvamp
    M, w, its = 1_000_000, 500, 3
    dev = torch.device("cuda", 0)
    U, ldb, band, r, x0, _ = bench.build_problem(torch, M, w, 5, dev)
    p = bench.vamp_params(M)
    probes = bench.make_probes(its, M, 5)

    def run(R):
        v = sgvamp.VAMP(N=bench.n_gwas(M), Nt=bench.n_gwas(M), M=M, K=1, rho=p["rho"], gamw=p["gamw"], gam1=p["gam1"],
                        a=np.array([1.0]), prior_vars=p["prior_vars"], prior_probs=p["prior_probs"], out_dir=None,
                        out_name="t")
        xs = v.infer(R, r, its, cg_maxit=p["cg_maxit"], em_prior_maxit=p["em_prior_maxit"], learn_gamw=True,
                     lmmse_damp=False, prior_update="em", update_prior_from=1, probes=probes)
        out = (xs, [v.history["rows"][i][0] for i in range(its)], [v.history["cg_iters"][i][0] for i in range(its)],
               v.handle.ld_info(0)["layout"])
        v.close()
        return out

    xs_a, rows_a, cg_a, lay_a = run(sgvamp.DeviceDSYM(U.data_ptr(), w, ldb, 0, keepalive=U))
    xs_b, rows_b, cg_b, lay_b = run(sgvamp.DeviceDSYM(U.data_ptr(), w, ldb, 0, keepalive=U))
    ldf = (M + 31) // 32 * 32
    full = torch.zeros((2 * w + 1, ldf), device=dev, dtype=torch.float32)
    full[:, :M] = band
    del band
    xs_c, rows_c, cg_c, lay_c = run(sgvamp.DeviceDIA(full.data_ptr(), w, ldf, keepalive=full))
    assert (lay_a, lay_c) == ("dsym", "dia")
    for it in range(its):
        assert np.array_equal(xs_a[it], xs_b[it]) and rows_a[it] == rows_b[it]            # bit-reproducible
        assert rel_l2(xs_a[it], xs_c[it]) <= 1e-9                                           # two kernel families agree
        assert rel_err(rows_a[it][1:6], rows_c[it][1:6]) <= 1e-9
        assert tuple(cg_a[it]) == tuple(cg_c[it])
        assert np.all(np.isfinite(xs_a[it]))
    al = float(np.dot(xs_a[-1].ravel(), x0) / (np.linalg.norm(xs_a[-1]) * np.linalg.norm(x0)))
    assert al > 0.5


# ---------------------------------------------------------------------------------------------
# multi-probe Hutchinson (extension): n_probes = 1 is the reference; n_probes > 1 is checked against the
# oracle extended in the same way (oracle/sgvamp_oracle.py: further probes solved from zero and averaged)
# ---------------------------------------------------------------------------------------------
def _run_oracle_probes(c, n_probes, probes):
    from oracle import sgvamp_oracle as orc
    K, M = c["K"], c["M"]
    o = orc.VAMPOracle(c["N_list"], M, c["rho"], c["gamw"], c["gam1"], c["prior_vars"], c["prior_probs"])
    Rs = [orc.regularise(R, c["s"]) for R in c["R"]]
    fn = lambda k, it, M_, p=0: probes[k, it, p]
    x0 = c["x0"] * np.sqrt(c["N_list"][0]) if "x0" in c else None
    return o.infer(Rs, list(c["r"]), c["iterations"], x0=x0, cg_maxit=c["cg_maxit"], em_prior_maxit=c["em_prior_maxit"],
                   learn_gamw=c["learn_gamw"], lmmse_damp=c["lmmse_damp"], prior_update=c["prior_update"],
                   update_prior_from=c["update_prior_from"], probe_fn=fn, n_probes=n_probes)


@pytest.mark.parametrize("name,n_probes", [("banded_L2_em_s01", 2), ("banded_L2_em_s01", 4), ("dense_K3_L2_em", 3),
                                           ("blockdiag_L3_em_s01", 5)])
def test_multi_probe_hutchinson_matches_oracle(nat, name, n_probes):
    import sgvamp
    c = load_case(name)
    K, M = c["K"], c["M"]
    rng = np.random.default_rng(n_probes)
    probes = (rng.integers(0, 2, size=(K, c["iterations"], n_probes, M)) * 2 - 1).astype(np.int8)
    probes[:, :, 0, :] = c["probes"]                         # probe 0 = the golden run's probe
    ref = _run_oracle_probes(c, n_probes, probes)
    N_list = c["N_list"]
    Nt = sum(N_list)
    v = sgvamp.VAMP(N=N_list if K > 1 else N_list[0], Nt=Nt, M=M, K=K, rho=c["rho"], gamw=c["gamw"], gam1=c["gam1"],
                    a=np.array(N_list) / Nt, prior_vars=c["prior_vars"], prior_probs=c["prior_probs"], out_dir=None, out_name="g")
    x0 = c["x0"] * np.sqrt(N_list[0]) if "x0" in c else None
    xs = v.infer(c["R"] if K > 1 else c["R"][0], list(c["r"]) if K > 1 else c["r"][0], c["iterations"], x0=x0,
                 cg_maxit=c["cg_maxit"], em_prior_maxit=c["em_prior_maxit"], learn_gamw=c["learn_gamw"], lmmse_damp=c["lmmse_damp"],
                 prior_update=c["prior_update"], update_prior_from=c["update_prior_from"], s=c["s"], probes=probes,
                 n_probes=n_probes)
    differs = False
    for it in range(c["iterations"]):
        # tolerance: the GPU holds LD in fp32 (the oracle in fp64), as in the golden trajectory tests
        assert rel_l2(xs[it], ref["xhat1"][it]) <= 1e-5, (it, rel_l2(xs[it], ref["xhat1"][it]))
        for k in range(K):
            assert rel_err(v.history["rows"][it][k][1:6], np.array(ref["rows"][it][k][1:6])) <= 1e-5
            assert tuple(v.history["cg_iters"][it][k]) == tuple(ref["cg_iters"][it][k])
            differs = differs or rel_err(v.history["rows"][it][k][5:6], c["rows"][it, k, 5:6]) > 1e-6
    assert differs                                            # the extra probes do change alpha2 relative to the 1-probe golden
    v.close()


def test_outputs_parse_like_reference_plot_script(nat):
    """scripts/plots.py:34-58 reads the cohort CSV with csv.reader(delimiter='\\t'), skips the header and takes
    int(row[0]), float(row[1..6]) = it, gamw, gam1, gam2, alpha1, alpha2, lam, and float(row[1]), float(row[2]) =
    alignment, l2 from the metrics CSV: the files this library writes must go through exactly that reader."""
    import csv
    c = load_case("dense_L2_em")
    with tempfile.TemporaryDirectory() as d:
        xs, hist, info, fin = run_gpu(c, out_dir=d)
        its, cols = [], [[] for _ in range(6)]
        with open(os.path.join(d, "g_cohort_1.csv"), mode="r") as f:
            rd = csv.reader(f, delimiter="\t")
            next(rd, None)
            for row in rd:
                its.append(int(row[0]))
                for j in range(6):
                    cols[j].append(float(row[1 + j]))
        assert max(its) + 1 == c["iterations"] and its == list(range(c["iterations"]))
        for it in range(c["iterations"]):
            assert [cols[j][it] for j in range(6)] == [float(v) for v in hist["rows"][it][0][1:7]]
        al, l2 = [], []
        with open(os.path.join(d, "g_metrics.csv"), mode="r") as f:
            rd = csv.reader(f, delimiter="\t")
            next(rd, None)
            for row in rd:
                al.append(float(row[1]))
                l2.append(float(row[2]))
        assert len(al) == c["iterations"] and all(0.0 < a <= 1.0 for a in al) and all(v >= 0.0 for v in l2)
