"""Host-side layout arithmetic that the GPU paths rely on, checked on the CPU: the tiled index of the symmetric
half-band layout (ldgen.dsym_tile vs the formula of include/sgvamp_b200.h), the DIA column-window builder of
bench.py against scipy's own DIA semantics, and the half-band construction of bench.build_problem's recipe."""
import numpy as np
import pytest
import scipy.sparse
import torch

import bench
import ldgen


def dsym_index(j, d, ngr):
    return ((j >> 7) * ngr + (d >> 2)) * 512 + (d & 3) * 128 + (j & 127)


@pytest.mark.parametrize("Dp,ldb", [(4, 128), (8, 256), (504, 384), (12, 1280)])
def test_dsym_tile_matches_index_formula(Dp, ldb):
    U = torch.arange(Dp * ldb, dtype=torch.float32).reshape(Dp, ldb)
    T = ldgen.dsym_tile(torch, U).reshape(-1).numpy()
    ngr = Dp // 4
    rng = np.random.default_rng(0)
    for _ in range(200):
        d, j = int(rng.integers(0, Dp)), int(rng.integers(0, ldb))
        assert T[dsym_index(j, d, ngr)] == U[d, j].item()
    assert T.size == Dp * ldb                                  # a permutation: nothing added, nothing lost
    assert np.array_equal(np.sort(T), np.arange(Dp * ldb, dtype=np.float32))


def _random_band(M, w, seed):
    rng = np.random.default_rng(seed)
    band = np.zeros((2 * w + 1, M), dtype=np.float32)          # band[k, i] = R[i, i + k - w]
    for k in range(2 * w + 1):
        off = k - w
        lo, hi = max(0, -off), min(M, M - off)
        band[k, lo:hi] = rng.standard_normal(hi - lo).astype(np.float32)
    return band


def _dense_from_band(band, M, w):
    R = np.zeros((M, M))
    for k in range(2 * w + 1):
        off = k - w
        for i in range(max(0, -off), min(M, M - off)):
            R[i, i + off] = band[k, i]
    return R


@pytest.mark.parametrize("M,w,glo,hi", [(40, 3, 0, 40), (64, 5, 16, 48), (64, 5, 0, 20), (50, 7, 30, 50), (30, 4, 10, 11)])
def test_band_to_host_dia_window_is_scipy_dia(M, w, glo, hi):
    """bench.band_to_host_dia: rows [glo, hi) of a band as a column window of scipy's DIA arrays."""
    band = _random_band(M, w, seed=M + w + glo)
    R = _dense_from_band(band, M, w)
    host, offsets, col0 = bench.band_to_host_dia(torch, torch.from_numpy(band[:, glo:hi].copy()), M, w, glo, hi, pinned=False)
    data = host.numpy()
    ldd = data.shape[1]
    assert col0 == max(0, glo - w) and col0 + ldd == min(M, hi + w)
    # scipy semantics: data[k, j] = A[j - off_k, j]; embed the window into full-width arrays and let scipy rebuild A
    full = np.zeros((2 * w + 1, M), dtype=np.float32)
    full[:, col0:col0 + ldd] = data
    A = scipy.sparse.dia_matrix((full, offsets), shape=(M, M)).toarray()
    assert np.array_equal(A[glo:hi], R[glo:hi].astype(np.float32))      # the rows of the window are exact
    other = np.ones(M, dtype=bool)
    other[glo:hi] = False
    assert not A[other].any()                                           # and nothing else is in it


def test_half_band_recipe_of_the_bench():
    """U[d, j] = R[i, i + d] (diagonal halved, extension rows keep only couplings to own rows), as bench.build_problem
    lays it out before tiling - checked against a dense symmetric matrix."""
    M, w, lo, hi, ext = 48, 5, 16, 40, 8
    band = _random_band(M, w, seed=3)
    for d in range(1, w + 1):                                   # symmetrise as the bench does
        band[w - d, d:] = band[w + d, : M - d]
    R = _dense_from_band(band, M, w)
    assert np.array_equal(R, R.T)
    glo, n = lo - ext, hi - (lo - ext)
    b = torch.from_numpy(band[:, glo:hi].copy())
    Dp = (w + 1 + 3) // 4 * 4
    U = torch.zeros((Dp, n), dtype=torch.float32)
    U[: w + 1] = b[w:]
    U[0] *= 0.5
    jj = torch.arange(ext)[None, :]
    dd = torch.arange(Dp)[:, None]
    U[:, :ext] *= (jj + dd >= ext).to(torch.float32)
    U = U.numpy()
    # product restricted to own rows from the half band: y[i] = sum_d U[d][i] x[i+d] + sum_d U[d][i-d] x[i-d]
    x = np.random.default_rng(1).standard_normal(M)
    y = np.zeros(M)
    for j in range(n):
        i = glo + j
        for d in range(Dp):
            if U[d, j] != 0.0 and i + d < M:
                if lo <= i < hi:
                    y[i] += U[d, j] * x[i + d]                  # forward
                if lo <= i + d < hi:
                    y[i + d] += U[d, j] * x[i]                  # transposed (d = 0: the second half of the diagonal)
    ref = R.astype(np.float32).astype(np.float64) @ x
    # rows whose right-hand couplings stay inside [glo, hi) are complete
    ok = np.arange(lo, hi - w)
    assert np.allclose(y[ok], ref[ok], rtol=1e-12, atol=1e-12)
