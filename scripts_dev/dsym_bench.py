"""Isolated timing of the DSYM / DIA SpMM kernels on a random band (no VAMP): ms per pass and GB/s."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "sgvamp-py_b200"))
import numpy as np, torch
import sgv_native as nat
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
w = int(sys.argv[2]) if len(sys.argv) > 2 else 500
Dp = (w + 1 + 3) // 4 * 4
ldb = (M + 127) // 128 * 128
U = torch.randn((Dp, ldb), device="cuda", dtype=torch.float32) * 0.01
U[w + 1:] = 0
for d in range(1, w + 1):
    U[d, M - d:] = 0
U[:, M:] = 0
import ldgen
U = ldgen.dsym_tile(torch, U)
h = nat.Handle()
h.configure(M, 1)
h.adopt_dsym(0, U.data_ptr(), w, ldb, 0)
info = h.ld_info(0)
x = np.random.default_rng(0).standard_normal((M, 2))
y = h.spmm(0, x)
for rep in range(3):
    ms = h.spmm_bench(0, 30)
    print("dsym M=%d w=%d: %.4f ms/pass  %.0f GB/s algorithmic" % (M, w, ms, info["bytes_per_pass"] / ms / 1e6))
h.close()
