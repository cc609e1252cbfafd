"""Seeded synthetic LD (R) and XTy (r) generators for tests and benchmarks.

The dense recipe is the reference's own (simulation/sim_gen_phen_mult.py:28-55):
X ~ Binomial(2, 0.4), column-standardised, sparse beta, y = X beta + noise, X /= sqrt(N),
r = X^T y, R = X^T X.  The block-diagonal and banded recipes produce genotype-derived LD with
the structure BASELINE.json's configs name (per-chromosome LD blocks; banded LD), PSD by
construction: genotypes come from thresholded latent Gaussian haplotypes with short-range
correlation; the banded matrix is the sample LD restricted to |i-j| <= w and multiplied by a
Bartlett taper (Schur product of PSD matrices stays PSD).

numpy functions here serve tests and golden vectors (small M).  The ``*_device`` functions use
torch on the GPU as a data-generation utility only (they are not part of the solver path).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse


# ------------------------------------------------------------------------------------------
# small / host generators
# ------------------------------------------------------------------------------------------
def sim_dense(M, N, lam=0.01, h2=0.5, seed=0, K=1, N_list=None):
    """Reference recipe.  Returns (R_list, r_list, beta, N_list); all cohorts share beta."""
    rng = np.random.default_rng(seed)
    cm = max(1, int(M * lam))
    beta = np.zeros(M)
    idx = rng.choice(M, cm, replace=False)
    beta[idx] = rng.normal(0, np.sqrt(h2 / cm), cm)
    N_list = list(N_list) if N_list is not None else [N] * K
    Rs, rs = [], []
    for Nk in N_list:
        X = rng.binomial(2, 0.4, size=(Nk, M)).astype(np.float64)
        X = (X - X.mean(axis=0)) / X.std(axis=0)
        y = X @ beta + rng.normal(0, np.sqrt(1 - h2), Nk)
        X /= np.sqrt(Nk)
        rs.append(X.T @ y)
        Rs.append(X.T @ X)
    return Rs, rs, beta, N_list


def _haplotype_genotypes(n, m, rng, rho=0.97):
    """n x m standardised genotypes from two thresholded latent AR(1) Gaussian haplotypes."""
    maf = rng.uniform(0.05, 0.5, m)
    from scipy.stats import norm as _norm
    thr = _norm.ppf(maf)
    G = np.zeros((n, m))
    for _ in range(2):
        e = rng.standard_normal((n, m))
        z = np.empty_like(e)
        z[:, 0] = e[:, 0]
        c = np.sqrt(1 - rho * rho)
        for j in range(1, m):
            z[:, j] = rho * z[:, j - 1] + c * e[:, j]
        G += (z < thr[None, :])
    sd = G.std(axis=0)
    sd[sd == 0] = 1.0
    return (G - G.mean(axis=0)) / sd


def bartlett_band(M, w):
    """Sparse Bartlett taper 1-|i-j|/(w+1) on |i-j| <= w (CSR)."""
    offs = np.arange(-w, w + 1)
    diags = [np.full(M - abs(o), 1.0 - abs(o) / (w + 1.0)) for o in offs]
    return scipy.sparse.diags(diags, offs, shape=(M, M), format="csr")


def sim_banded(M, w, N_ld=512, N=None, lam=0.01, h2=0.5, seed=0, exact_noise=False):
    """Banded LD (CSR, |i-j|<=w, unit diagonal) + r.  Returns (R, r, x0_scaled, N)."""
    rng = np.random.default_rng(seed)
    N = N if N is not None else N_ld
    X = _haplotype_genotypes(N_ld, M, rng) / np.sqrt(N_ld)
    # banded sample LD, chunked so that M x M is never formed
    rows, cols, vals = [], [], []
    step = max(w, 256)
    for lo in range(0, M, step):
        hi = min(M, lo + step)
        clo, chi = max(0, lo - w), min(M, hi + w)
        B = X[:, lo:hi].T @ X[:, clo:chi]
        ii, jj = np.meshgrid(np.arange(lo, hi), np.arange(clo, chi), indexing="ij")
        keep = np.abs(ii - jj) <= w
        taper = 1.0 - np.abs(ii - jj)[keep] / (w + 1.0)
        rows.append(ii[keep]); cols.append(jj[keep]); vals.append(B[keep] * taper)
    R = scipy.sparse.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(M, M))
    R.setdiag(1.0)
    R.sort_indices()
    cm = max(1, int(M * lam))
    beta = np.zeros(M)
    beta[rng.choice(M, cm, replace=False)] = rng.normal(0, np.sqrt(h2 / cm), cm)
    x0 = beta * np.sqrt(N)
    if exact_noise:
        # noise ~ N(0, R) exactly: Bartlett kernel = sum of boxcar outer products (see banded_dia_device)
        E = rng.standard_normal((N_ld, M + w))
        C = np.cumsum(E, axis=1)
        F = C[:, w:].copy()
        F[:, 1:] -= C[:, :M - 1]
        noise = (X * F).sum(axis=0) / np.sqrt(w + 1.0)
    else:
        noise = X.T @ rng.standard_normal(N_ld)
    r = R @ x0 + np.sqrt(1 - h2) * noise
    return R, r, x0, N


def sim_blockdiag(M, block_lo=100, block_hi=400, N_ld=512, N=None, lam=0.01, h2=0.5, seed=0):
    """Block-diagonal LD (dense inside LD blocks, CSR).  Returns (R, r, x0_scaled, N, block_starts)."""
    rng = np.random.default_rng(seed)
    N = N if N is not None else N_ld
    sizes = []
    left = M
    while left > 0:
        b = int(min(left, rng.integers(block_lo, block_hi)))
        sizes.append(b)
        left -= b
    starts = np.concatenate([[0], np.cumsum(sizes)])
    blocks, Xs = [], []
    for b in sizes:
        Xb = _haplotype_genotypes(N_ld, b, rng) / np.sqrt(N_ld)
        Rb = Xb.T @ Xb
        np.fill_diagonal(Rb, 1.0)
        blocks.append(Rb)
        Xs.append(Xb)
    R = scipy.sparse.block_diag(blocks, format="csr")
    R.sort_indices()
    X = np.concatenate(Xs, axis=1)
    cm = max(1, int(M * lam))
    beta = np.zeros(M)
    beta[rng.choice(M, cm, replace=False)] = rng.normal(0, np.sqrt(h2 / cm), cm)
    x0 = beta * np.sqrt(N)
    r = R @ x0 + np.sqrt(1 - h2) * (X.T @ rng.standard_normal(N_ld))
    return R, r, x0, N, starts


def round_to_f32(R):
    """Round matrix values to fp32-representable fp64 (so fp32 device storage is exact)."""
    if scipy.sparse.issparse(R):
        R = R.copy()
        R.data = R.data.astype(np.float32).astype(np.float64)
        return R
    return np.asarray(R).astype(np.float32).astype(np.float64)


# ------------------------------------------------------------------------------------------
# device generators (torch as a data-generation utility; used by bench.py and big GPU tests)
# ------------------------------------------------------------------------------------------
def _chunked_randn(torch, n, lo, hi, seed, tag, dev, chunk=8192):
    """n x (hi-lo) standard normals whose column j depends only on (seed, tag, j): any index range
    (negative indices included) can be generated independently on any rank."""
    c0, c1 = lo // chunk, (hi - 1) // chunk
    parts = []
    for c in range(c0, c1 + 1):
        g = torch.Generator(device=dev)
        g.manual_seed((seed * 1000003 + tag * 7919 + (c + 65536) * 104729) & 0x7FFFFFFFFFFF)
        parts.append(torch.randn((n, chunk), generator=g, device=dev, dtype=torch.float32))
    return torch.cat(parts, dim=1)[:, lo - c0 * chunk: hi - c0 * chunk]


def _latent_block(torch, n, lo, hi, seed, hap, B, kern, dev):
    """Latent Gaussian field z[:, lo:hi] = sum_i kern[j-i] eps_i (causal FIR filter of white noise,
    applied as a Toeplitz matmul)."""
    E = _chunked_randn(torch, n, lo - B, hi, seed, hap, dev)          # n x (hi-lo+B)
    m = hi - lo
    z = torch.empty((n, m), device=dev, dtype=torch.float32)
    T = 2048
    idx = torch.arange(T, device=dev)
    d = (idx[None, :] + B) - torch.arange(T + B, device=dev)[:, None]   # Toep[i_in, j_out] = kern[j_out + B - i_in]
    Toep = torch.where((d >= 0) & (d <= B), kern[d.clamp(0, B)], torch.zeros((), device=dev))
    for s in range(0, m, T):
        t = min(T, m - s)
        z[:, s:s + t] = E[:, s:s + t + B] @ Toep[:t + B, :t]
    return z


def genotype_counts_device(torch, n, lo, hi, seed, dev, rho=0.97, B=256):
    """Genotype counts in {0,1,2} of the markers [lo, hi) for n samples (n x (hi-lo), fp32): two thresholded latent
    AR(1)-like Gaussian haplotypes with a per-marker allele frequency; depends only on (seed, marker index)."""
    kern = (rho ** torch.arange(B + 1, device=dev, dtype=torch.float32))
    kern = kern / kern.norm()
    j = torch.arange(lo, hi, device=dev, dtype=torch.float64)
    # deterministic per-marker MAF in [0.05, 0.5] from a hash of the marker index
    u = torch.frac(torch.sin(j * 12.9898 + seed * 78.233) * 43758.5453).abs()
    maf = (0.05 + 0.45 * u).to(torch.float32)
    thr = torch.special.ndtri(maf.to(torch.float64)).to(torch.float32)
    G = torch.zeros((n, hi - lo), device=dev, dtype=torch.float32)
    for hap in range(2):
        z = _latent_block(torch, n, lo, hi, seed, hap, B, kern, dev)
        G += (z < thr[None, :]).to(torch.float32)
    return G


def genotypes_device(torch, n, lo, hi, seed, dev, rho=0.97, B=256):
    """Standardised genotype columns [lo, hi) (n x (hi-lo), fp32, already / sqrt(n))."""
    G = genotype_counts_device(torch, n, lo, hi, seed, dev, rho, B)
    mu = G.mean(dim=0, keepdim=True)
    sd = G.std(dim=0, unbiased=False, keepdim=True)
    sd = torch.where(sd == 0, torch.ones_like(sd), sd)
    return (G - mu) / sd / float(np.sqrt(n))


def banded_dia_device(torch, M, w, lo, hi, seed, dev, N_ld=4096, chunk=4096):
    """Rows [lo, hi) of the Bartlett-tapered banded sample LD in diagonal-major (DIA) layout.

    Returns (band, noise): fp32 band[d, i-lo] = R[i, i+d-w] (zero outside the matrix, unit
    diagonal), shape (2w+1, hi-lo), and an fp64 vector noise ~ N(0, R) restricted to [lo, hi).
    R = T o (X^T X) with the Bartlett kernel T[i,j] = max(0, 1-|i-j|/(w+1)) = sum_c t_c t_c^T,
    t_c = 1[c <= . <= c+w]/sqrt(w+1); hence noise_i = x_i . f_i with f_i = sum_{c=i-w..i} e_c /
    sqrt(w+1) for iid e_c ~ N(0, I) has covariance exactly R.  Every quantity depends only on
    (seed, marker index), so ranks can generate their row ranges independently.
    """
    nloc = hi - lo
    band = torch.zeros((2 * w + 1, nloc), device=dev, dtype=torch.float32)
    noise = torch.empty((nloc,), device=dev, dtype=torch.float64)
    dd = torch.arange(2 * w + 1, device=dev)
    taper = (1.0 - (dd - w).abs().to(torch.float32) / (w + 1.0))
    for s in range(lo, hi, chunk):
        t = min(hi, s + chunk)
        clo, chi = max(0, s - w), min(M, t + w)
        X = genotypes_device(torch, N_ld, clo, chi, seed, dev)
        Xi = X[:, s - clo: t - clo]
        P = Xi.T @ X                                   # (t-s) x (chi-clo), fp32
        E = _chunked_randn(torch, N_ld, s - w, t, seed, 5, dev)          # columns c = s-w .. t-1
        C = torch.cumsum(E.to(torch.float64), dim=1)
        F = C[:, w:].clone()                                             # sum_{c <= i}
        F[:, 1:] -= C[:, : t - s - 1]                                    # minus sum_{c < i-w}
        noise[s - lo: t - lo] = (Xi.to(torch.float64) * F).sum(dim=0) / float(np.sqrt(w + 1.0))
        ii = torch.arange(s, t, device=dev)
        col = ii[None, :] + (dd[:, None] - w)          # (2w+1) x (t-s): global column
        ok = (col >= 0) & (col < M)
        cidx = (col - clo).clamp(0, chi - clo - 1)
        vals = P[(ii - s)[None, :].expand_as(cidx), cidx] * taper[:, None]
        vals = torch.where(ok, vals, torch.zeros((), device=dev))
        band[:, s - lo: t - lo] = vals
        del X, P, E, C, F
    band[w, :] = 1.0
    return band, noise


def banded_dsym_library(torch, nat, M, w, lo, hi, rank, world, seed, dev, s, N_ld=4096, chunk=8192):
    """Rows [lo, hi) of the same workload as banded_dia_device, with the LD matrix built by the LIBRARY's kernel
    (sgv_ld_build_banded: integer Gram sums of the int8 genotypes, standardisation, Bartlett taper and Rused = (1-s) R + s I
    in its epilogue, written straight into the tiled half-band layout) instead of torch matmuls.  torch only synthesises
    the genotypes and the exact N(0, R) noise.  Returns (U tiled fp32 tensor, ldb, ext, noise fp64 for rows [lo, hi))."""
    h = nat.Handle(device=dev.index or 0)
    if world > 1:
        h.configure_part(M, 1, rank, world, lo, hi, True)
    else:
        h.configure(M, 1)
    ext = h.dsym_extension(w)
    g0, g1 = max(0, lo - ext), min(M, hi + w)
    ldg = (N_ld + 15) // 16 * 16
    Gt = torch.zeros((g1 - g0, ldg), device=dev, dtype=torch.int8)
    noise = torch.empty((hi - lo,), device=dev, dtype=torch.float64)
    for c0 in range(g0, g1, chunk):
        c1 = min(g1, c0 + chunk)
        G = genotype_counts_device(torch, N_ld, c0, c1, seed, dev)            # N_ld x (c1-c0)
        Gt[c0 - g0: c1 - g0, :N_ld] = G.t().to(torch.int8)
        a0, a1 = max(c0, lo), min(c1, hi)                                      # own rows of this chunk: their noise
        if a1 > a0:
            Gi = G[:, a0 - c0: a1 - c0].to(torch.float64)
            mu = Gi.mean(dim=0, keepdim=True)
            sd = Gi.std(dim=0, unbiased=False, keepdim=True)
            sd = torch.where(sd == 0, torch.ones_like(sd), sd)
            Xi = (Gi - mu) / sd / float(np.sqrt(N_ld))
            E = _chunked_randn(torch, N_ld, a0 - w, a1, seed, 5, dev)
            Cs = torch.cumsum(E.to(torch.float64), dim=1)
            F = Cs[:, w:].clone()
            F[:, 1:] -= Cs[:, : a1 - a0 - 1]
            noise[a0 - lo: a1 - lo] = (Xi * F).sum(dim=0) / float(np.sqrt(w + 1.0))
            del Gi, Xi, E, Cs, F
        del G
    h.build_banded(0, None, N_ld, w, s=s, taper=True, g0=g0, device_ptr=Gt.data_ptr(), nmark=g1 - g0, ldg=ldg)
    _w, ldb, ext2 = h.band_shape(0)
    assert ext2 == ext
    Dp = (w + 1 + 3) // 4 * 4
    U = torch.empty((Dp * ldb,), device=dev, dtype=torch.float32)
    h.copy_band(0, U.data_ptr(), U.numel())
    h.close()
    del Gt
    return U, ldb, ext, noise


def dsym_untile(torch, U, Dp, ldb):
    """Inverse of dsym_tile: the tiled buffer -> (Dp, ldb) diagonal-major half band."""
    return U.view(ldb // 128, Dp // 4, 4, 128).permute(1, 2, 0, 3).reshape(Dp, ldb)


def dsym_tile(torch, U):
    """(Dp, ldb) diagonal-major half band (Dp % 4 == 0, ldb % 128 == 0) -> the tiled DSYM layout of
    sgv_ld_adopt_dsym: [ldb/128][Dp/4][4][128], contiguous."""
    Dp, ldb = U.shape
    assert Dp % 4 == 0 and ldb % 128 == 0
    return U.view(Dp // 4, 4, ldb // 128, 128).permute(2, 0, 1, 3).contiguous()


def causal_effects(M, N, lam, h2, seed):
    """Sparse true effects scaled by sqrt(N) (the x0 the solver estimates)."""
    rng = np.random.default_rng(seed + 99)
    cm = max(1, int(M * lam))
    beta = np.zeros(M)
    beta[rng.choice(M, cm, replace=False)] = rng.normal(0, np.sqrt(h2 / cm), cm)
    return beta * np.sqrt(N)
