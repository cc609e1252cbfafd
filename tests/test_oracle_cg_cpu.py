"""The oracle restates `scipy.sparse.linalg.cg` (third-party, unpinned by the reference: src/sgvamp.py:7,316,332);
this pins the restatement to the scipy that is installed here: same iterates, same `info`, same number of matvecs,
for cold and warm starts, zero right-hand sides, exhausted and sufficient iteration budgets."""
import numpy as np
import pytest
import scipy.sparse
import scipy.sparse.linalg
from hypothesis import given, settings, strategies as st

from oracle import sgvamp_oracle as orc


def _spd(n, seed, cond):
    rng = np.random.default_rng(seed)
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    ev = np.geomspace(1.0, cond, n)
    return (Q * ev) @ Q.T


@settings(max_examples=40, deadline=None)
@given(n=st.integers(2, 60), seed=st.integers(0, 10_000), cond=st.sampled_from([1.5, 30.0, 1e3, 1e6]),
       maxiter=st.sampled_from([0, 1, 2, 5, 50, 500]), warm=st.booleans(), zero_b=st.booleans())
def test_cg_restates_scipy(n, seed, cond, maxiter, warm, zero_b):
    A = _spd(n, seed, cond)
    rng = np.random.default_rng(seed + 1)
    b = np.zeros(n) if zero_b else rng.standard_normal(n)
    x0 = rng.standard_normal(n) if warm else np.zeros(n)
    calls = [0]

    def mv(v):
        calls[0] += 1
        return A @ v

    op = scipy.sparse.linalg.LinearOperator((n, n), matvec=mv, dtype=np.float64)
    xs, info_s = scipy.sparse.linalg.cg(op, b, maxiter=maxiter, x0=x0.copy())
    n_scipy = calls[0]
    calls[0] = 0
    xo, info_o, n_updates = orc.cg(mv, b, x0.copy(), maxiter)
    assert info_o == info_s
    assert np.array_equal(xo, xs)                    # the same floating-point operations in the same order
    assert calls[0] == n_scipy                       # including the residual matvec of a warm start
    assert n_updates == n_scipy - (1 if (warm and not zero_b and x0.any()) else 0)


def test_cg_matches_scipy_on_sparse_banded():
    M = 400
    d = [np.full(M - abs(o), 0.3 ** abs(o)) for o in range(-3, 4)]
    R = scipy.sparse.diags(d, list(range(-3, 4)), format="csr")
    A = (2.5 * R + 0.8 * scipy.sparse.identity(M)).tocsr()
    b = np.random.default_rng(0).standard_normal(M)
    xs, info_s = scipy.sparse.linalg.cg(A, b, maxiter=500)
    xo, info_o, _ = orc.cg(lambda v: A @ v, b, np.zeros(M), 500)
    assert info_s == info_o == 0 and np.array_equal(xs, xo)
