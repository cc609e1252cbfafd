"""Isolated timing of the general-CSR SpMM kernel (k_spmm_csr) on an irregular symmetric sparse matrix: ms per 2-RHS pass,
algorithmic GB/s (8 B per stored value + row pointers + vector pair in/out)."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "sgvamp-py_b200"))
import numpy as np, scipy.sparse
import sgv_native as nat
M = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
nd = int(sys.argv[2]) if len(sys.argv) > 2 else 64
rng = np.random.default_rng(0)
offs = np.sort(rng.choice(np.arange(1, 20000), nd, replace=False))
diags = [np.ones(M)] + [rng.standard_normal(M - o).astype(np.float32).astype(np.float64) * 0.01 for o in offs]
U = scipy.sparse.diags(diags, [0] + list(offs), shape=(M, M), format="csr")
R = (U + scipy.sparse.triu(U, 1).T).tocsr()
R.sort_indices()
h = nat.Handle()
h.configure(M, 1)
h._ck(h.upload_csr(0, R.indptr, R.indices, R.data, layout=nat.LAYOUT_CSR))
info = h.ld_info(0)
x = rng.standard_normal((M, 2))
y = h.spmm(0, x)
print("csr check rel err %.2e" % (np.linalg.norm(y - R @ x) / np.linalg.norm(R @ x)))
for rep in range(3):
    ms = h.spmm_bench(0, 30)
    print("csr M=%d nnz=%d (%.0f per row): %.4f ms/pass  %.0f GB/s algorithmic" % (M, R.nnz, R.nnz / M, ms, info["bytes_per_pass"] / ms / 1e6))
h.close()
