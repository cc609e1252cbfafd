"""Tuning harness: time the DIA SpMM (EPI_Q) for one SGV_DIA_CFG variant on a random band."""
import sys, os
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "sgvamp-py_b200"))
import torch, sgv_native as nat
M = int(sys.argv[1]); w = int(sys.argv[2])
ldb = (M + 31) // 32 * 32
band = torch.randn((2 * w + 1, ldb), device="cuda", dtype=torch.float32)
h = nat.Handle(); h.configure(M, 1)
h.adopt_dia(0, band.data_ptr(), w, ldb)
x = np.random.default_rng(0).standard_normal((M, 2))
y = h.spmm(0, x)            # fills pp with x (EPI_PLAIN path), sanity value below
ms = min(h.spmm_bench(0, 20) for _ in range(3))
info = h.ld_info(0)
print("cfg %s M=%d w=%d: %.4f ms  %.0f GB/s  (%.1f%% of 6550)  chk=%.6e" % (
    os.environ.get("SGV_DIA_CFG", "default"), M, w, ms, info["bytes_per_pass"] / ms / 1e6, info["bytes_per_pass"] / ms / 1e6 / 65.5, float(np.abs(y).sum())))
