"""Function-level pinning of the oracle against the UNMODIFIED reference module (imported from /root/reference when
it is present - the build container; skipped elsewhere, where the committed goldens carry the pinning):
denoiser_meta / der_denoiser_meta, prior_update_em, Lagrangian_der on random inputs for random K and L."""
import os
import sys
import types

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import sgvamp_oracle as orc

REF = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "sgvamp.py")), reason="reference tree not present")


class _Comm:
    def __init__(self, rank=0, size=1):
        self.rank, self.size = rank, size

    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.size

    def bcast(self, x, root=0):
        return x


def _ref_module():
    if "ref_sgvamp" in sys.modules:
        return sys.modules["ref_sgvamp"]
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_sgvamp", os.path.join(REF, "sgvamp.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_sgvamp"] = mod
    spec.loader.exec_module(mod)
    return mod


def _make(K, L, rank, seed, tmp, M=9):
    rng = np.random.default_rng(seed)
    probs = rng.dirichlet(np.ones(L))
    probs[0] = 0.5 + 0.5 * probs[0]
    probs[1:] *= (1 - probs[0]) / probs[1:].sum()
    pvars = np.concatenate([[0.0], np.sort(rng.uniform(1e-4, 1e-2, L - 1))])
    Ns = rng.integers(500, 3000, K)
    a = Ns / Ns.sum()
    ref = _ref_module().VAMP(N=int(Ns[rank]), Nt=int(Ns.sum()), M=M, K=K, rho=0.5, gamw=2.0, gam1=1e-3, a=a,
                             prior_vars=list(pvars), prior_probs=list(probs), out_dir=tmp, out_name="f",
                             comm=_Comm(rank, K))
    prior = orc.Prior(list(pvars), list(probs), int(Ns.sum()))
    return ref, prior, a, rng


@settings(max_examples=60, deadline=None)
@given(K=st.integers(1, 4), L=st.integers(2, 6), seed=st.integers(0, 100000))
def test_denoiser_and_derivative(K, L, seed, tmp_path_factory):
    tmp = str(tmp_path_factory.mktemp("d"))
    rank = seed % K
    ref, prior, a, rng = _make(K, L, rank, seed, tmp)
    M = 9
    r1s = rng.standard_normal((K, M)) * rng.choice([0.1, 1.0, 10.0])
    gam1s = rng.uniform(1e-3, 5.0, K)
    xhat, dfac = orc.denoise_all(r1s, gam1s, a, prior)
    xh1, df1 = orc.denoise_all(r1s, gam1s, a, prior, per_marker=True)
    for j in range(M):
        want = ref.denoiser_meta(r1s[:, j], gam1s)
        wder = ref.der_denoiser_meta(r1s[:, j], gam1s)
        assert np.isclose(xhat[j], want, rtol=1e-10, atol=1e-300)
        assert xh1[j] == want
        # literal form with the rank factor inside the sums, where the reference has it (:112-113): bit-identical
        w = a * gam1s
        lit = orc._denoise_one(r1s[:, j], w, prior.lam, prior.omegas, prior.sigmas, w_rank=(a[rank], gam1s[rank]))
        assert lit[0] == want and lit[1] == wder
        # factored form (what the kernels compute: a[rank]*gam1[rank] outside the sums): the derivative is a difference
        # of two products (:114), so the bound scales with the cancelled magnitude DerDen*Num/Den^2 = xhat^2
        scale = abs(wder) + a[rank] * gam1s[rank] * want * want
        assert abs(a[rank] * gam1s[rank] * dfac[j] - wder) <= 1e-10 * scale + 1e-300
        assert abs(a[rank] * gam1s[rank] * df1[j] - wder) <= 1e-12 * scale + 1e-300


@settings(max_examples=40, deadline=None)
@given(K=st.integers(1, 3), L=st.integers(2, 5), seed=st.integers(0, 100000))
def test_em_pass_and_lagrangian(K, L, seed, tmp_path_factory):
    tmp = str(tmp_path_factory.mktemp("e"))
    M = 50
    ref, prior, a, rng = _make(K, L, 0, seed, tmp, M)
    r1s = rng.standard_normal((K, M)) * 3.0
    gam1s = rng.uniform(1e-2, 3.0, K)
    ref.prior_update_em(r1s, gam1s)
    orc.prior_update_em(r1s, gam1s, a, prior)
    assert np.isclose(prior.lam, ref.lam, rtol=1e-12)
    assert np.allclose(prior.omegas, ref.omegas, rtol=1e-12)
    omega0 = np.concatenate([[1 - prior.lam], prior.lam * prior.omegas])
    sigma2 = np.concatenate([[1e-16], prior.sigmas])
    x = np.concatenate([omega0 * rng.uniform(0.8, 1.2, L), [rng.uniform(0.5, 2.0)]])
    want = ref.Lagrangian_der(x, omega0, sigma2, r1s, gam1s)
    got = orc.lagrangian_der(x, omega0, sigma2, r1s, gam1s, a, L)
    assert np.allclose(got, want, rtol=1e-11, atol=1e-9)
