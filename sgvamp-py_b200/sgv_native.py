"""ctypes binding of libsgvamp_b200.so (C ABI declared in include/sgvamp_b200.h).

There is no CPU fallback: if the shared library is missing or no B200 is present, loading /
creating a handle raises.  Nothing in this module imports the oracle.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SGV_LIB") or os.path.join(HERE, "libsgvamp_b200.so")   # SGV_LIB: A/B of builds (dev only)

LAYOUT_AUTO, LAYOUT_DENSE, LAYOUT_DIA, LAYOUT_BLOCKDIAG, LAYOUT_CSR, LAYOUT_DSYM = 0, 1, 2, 3, 4, 5
LAYOUT_NAMES = {0: "auto", 1: "dense", 2: "dia", 3: "blockdiag", 4: "csr", 5: "dsym"}
F32, F64 = 0, 1
VEC_XHAT1, VEC_R1, VEC_XHAT2, VEC_SIGMA2U, VEC_R2, VEC_XTY = 0, 1, 2, 3, 4, 5

# every symbol the header declares (tests check that the library exports all of them)
SYMBOLS = [
    "sgv_version", "sgv_last_error", "sgv_create", "sgv_destroy", "sgv_sync", "sgv_configure",
    "sgv_ld_upload_dense", "sgv_ld_upload_csr", "sgv_ld_adopt_dia", "sgv_ld_adopt_dense", "sgv_ld_info",
    "sgv_set_xty", "sgv_reset_state", "sgv_get_vec", "sgv_set_vec", "sgv_get_vec_async", "sgv_wait_copies",
    "sgv_pinned_alloc", "sgv_pinned_free", "sgv_set_prior", "sgv_set_weights", "sgv_denoise", "sgv_prior_em",
    "sgv_lagrangian", "sgv_lmmse", "sgv_update_r1", "sgv_metrics", "sgv_spmm", "sgv_spmm_bench",
    "sgv_launch_count", "sgv_profile", "sgv_profile_read", "sgv_configure_part", "sgv_ipc_export", "sgv_ipc_import",
    "sgv_peer_attach_local", "sgv_partition_info", "sgv_ld_set_bandwidth_hint", "sgv_spmm_stage", "sgv_spmm_run",
    "sgv_ld_adopt_dsym", "sgv_dsym_extension", "sgv_set_host_barrier", "sgv_device_id", "sgv_ld_upload_dia",
    "sgv_iteration_supported", "sgv_vamp_begin", "sgv_set_truth", "sgv_iteration_probe_buffer", "sgv_iteration_enqueue",
    "sgv_iteration_wait", "sgv_ld_adopt_blockdiag", "sgv_ld_build_banded", "sgv_ld_copy_band", "sgv_vamp_set_alphas",
    "sgv_ld_upload_dense_rows", "sgv_ld_adopt_dense_colpanel", "sgv_r1_block", "sgv_probe_pair",
]
MAX_K, MAX_L, ITER_SLOTS = 8, 8, 4


class LmmseIn(C.Structure):
    _fields_ = [("gamw", C.c_double), ("gam2", C.c_double), ("alpha1", C.c_double), ("rho", C.c_double),
                ("cg_maxit", C.c_int), ("lmmse_damp", C.c_int), ("learn_gamw", C.c_int), ("x0_zero", C.c_int)]


class LmmseOut(C.Structure):
    _fields_ = [("u_sigma2u", C.c_double), ("xhat2_r", C.c_double), ("xhat2_R_xhat2", C.c_double),
                ("u_R_sigma2u", C.c_double), ("cg_iters", C.c_int * 2), ("cg_info", C.c_int * 2),
                ("spmm_passes", C.c_int)]


class ProbeOut(C.Structure):
    _fields_ = [("u_s", C.c_double * 2), ("u_R_s", C.c_double * 2), ("cg_iters", C.c_int * 2), ("cg_info", C.c_int * 2),
                ("spmm_passes", C.c_int)]


class IterIn(C.Structure):
    _fields_ = [("it", C.c_int), ("update_prior", C.c_int), ("em_maxit", C.c_int), ("em_tol", C.c_double),
                ("rho", C.c_double), ("cg_maxit", C.c_int), ("lmmse_damp", C.c_int), ("learn_gamw", C.c_int),
                ("want_metrics", C.c_int)]


class IterCohort(C.Structure):
    _fields_ = [("row", C.c_double * 7), ("cg_iters", C.c_int * 2), ("cg_info", C.c_int * 2), ("spmm_passes", C.c_int)]


class IterOut(C.Structure):
    _fields_ = [("lam", C.c_double), ("omegas", C.c_double * MAX_L), ("em_relerr", C.c_double), ("metrics", C.c_double * 4),
                ("em_steps", C.c_int), ("coh", IterCohort * MAX_K)]


class SgvError(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library (raises if it has not been built: no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SgvError("%s not found - build it with `python sgvamp-py_b200/build_native.py` "
                       "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.sgv_last_error.restype = C.c_char_p
    lib.sgv_launch_count.restype = C.c_int64
    lib.sgv_launch_count.argtypes = [C.c_void_p]
    _lib = lib
    return lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Handle:
    """One GPU's worth of solver state (opaque sgv_handle)."""

    def __init__(self, device=0, stream=None):
        self.lib = load()
        self.h = C.c_void_p()
        self._ck(self.lib.sgv_create(C.c_int(device), C.c_void_p(stream or 0), C.byref(self.h)))
        self.M = 0
        self.K = 0
        self._pinned = []

    def _ck(self, rc):
        if rc != 0:
            raise SgvError("libsgvamp_b200 error %d: %s" % (rc, self.lib.sgv_last_error().decode()))

    def close(self):
        if self.h:
            for p in self._pinned:
                self.lib.sgv_pinned_free(self.h, C.c_void_p(p))
            self._pinned = []
            self.lib.sgv_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- setup ---------------------------------------------------------------------------------
    def configure(self, M, K):
        self._ck(self.lib.sgv_configure(self.h, C.c_int64(M), C.c_int(K)))
        self.M, self.K = int(M), int(K)          # self.M: length of this handle's (local) vectors
        self.M_global, self.row_lo, self.rank, self.world = int(M), 0, 0, 1

    def configure_part(self, M, K, rank, world, row_lo, row_hi, halo):
        """Row partition: this handle owns marker rows [row_lo, row_hi) of an M-marker problem."""
        self._ck(self.lib.sgv_configure_part(self.h, C.c_int64(M), C.c_int(K), C.c_int(rank), C.c_int(world),
                                             C.c_int64(row_lo), C.c_int64(row_hi), C.c_int(int(halo))))
        self.M, self.K = int(row_hi - row_lo), int(K)
        self.M_global, self.row_lo, self.rank, self.world = int(M), int(row_lo), int(rank), int(world)

    def ipc_export(self):
        buf = C.create_string_buffer(64)
        self._ck(self.lib.sgv_ipc_export(self.h, buf))
        return bytes(buf.raw)

    def ipc_import(self, peer_rank, handle_bytes, peer_rows):
        self._ck(self.lib.sgv_ipc_import(self.h, C.c_int(peer_rank), C.c_char_p(handle_bytes), C.c_int64(peer_rows)))

    def peer_attach_local(self, peer_rank, other):
        self._ck(self.lib.sgv_peer_attach_local(self.h, C.c_int(peer_rank), other.h))

    def set_host_barrier(self, enable):
        self._ck(self.lib.sgv_set_host_barrier(self.h, C.c_int(int(enable))))

    def device_id(self):
        buf = C.create_string_buffer(32)
        self._ck(self.lib.sgv_device_id(self.h, buf, C.c_int(32)))
        return buf.value.decode()

    def set_bandwidth_hint(self, w):
        self._ck(self.lib.sgv_ld_set_bandwidth_hint(self.h, C.c_int64(int(w))))

    def sync(self):
        self._ck(self.lib.sgv_sync(self.h))

    def upload_dense(self, cohort, R, s=0.0):
        R = np.asarray(R)
        if R.dtype not in (np.float32, np.float64):
            R = R.astype(np.float64)
        if not R.flags.c_contiguous:
            R = np.ascontiguousarray(R)
        assert R.shape == (self.M_global, self.M_global), R.shape
        self._ck(self.lib.sgv_ld_upload_dense(self.h, C.c_int(cohort), R.ctypes.data_as(C.c_void_p),
                                              C.c_int(F64 if R.dtype == np.float64 else F32),
                                              C.c_int64(R.shape[1]), C.c_double(s)))

    def upload_dense_rows(self, cohort, rows, s=0.0):
        """Rows [row_lo, row_hi) of a dense R (this rank's slice of the rows partition, see sgv_ld_upload_dense_rows)."""
        rows = np.asarray(rows)
        if rows.dtype not in (np.float32, np.float64):
            rows = rows.astype(np.float64)
        if not rows.flags.c_contiguous:
            rows = np.ascontiguousarray(rows)
        assert rows.ndim == 2 and rows.shape[1] == self.M_global, rows.shape
        self._ck(self.lib.sgv_ld_upload_dense_rows(self.h, C.c_int(cohort), rows.ctypes.data_as(C.c_void_p),
                                                   C.c_int(F64 if rows.dtype == np.float64 else F32),
                                                   C.c_int64(rows.shape[1]), C.c_double(s)))

    def adopt_dense_colpanel(self, cohort, ptr, ld):
        self._ck(self.lib.sgv_ld_adopt_dense_colpanel(self.h, C.c_int(cohort), C.c_void_p(ptr), C.c_int64(ld)))

    def upload_csr(self, cohort, indptr, indices, data, s=0.0, layout=LAYOUT_AUTO):
        indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        indices = np.ascontiguousarray(indices, dtype=np.int32)
        data = np.ascontiguousarray(data)
        if data.dtype not in (np.float32, np.float64):
            data = data.astype(np.float64)
        rc = self.lib.sgv_ld_upload_csr(self.h, C.c_int(cohort), indptr.ctypes.data_as(C.POINTER(C.c_int64)),
                                        indices.ctypes.data_as(C.POINTER(C.c_int32)),
                                        data.ctypes.data_as(C.c_void_p),
                                        C.c_int(F64 if data.dtype == np.float64 else F32),
                                        C.c_int64(data.shape[0]), C.c_double(s), C.c_int(layout))
        return rc

    def upload_dia(self, cohort, data, offsets, s=0.0, layout=LAYOUT_AUTO, assume_symmetric=False, col0=0):
        """scipy DIA arrays: data[k, j - col0] = R[j - offsets[k], j] (see sgv_ld_upload_dia)."""
        data = np.ascontiguousarray(data)
        if data.dtype not in (np.float32, np.float64):
            data = data.astype(np.float64)
        offs = np.ascontiguousarray(offsets, dtype=np.int64)
        assert data.ndim == 2 and data.shape[0] == offs.shape[0]
        return self.lib.sgv_ld_upload_dia(self.h, C.c_int(cohort), data.ctypes.data_as(C.c_void_p),
                                          C.c_int(F64 if data.dtype == np.float64 else F32), C.c_int64(data.shape[1]),
                                          C.c_int64(col0), offs.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int(len(offs)),
                                          C.c_double(s), C.c_int(layout), C.c_int(int(bool(assume_symmetric))))

    def adopt_dia(self, cohort, band_ptr, w, ldb):
        self._ck(self.lib.sgv_ld_adopt_dia(self.h, C.c_int(cohort), C.c_void_p(band_ptr), C.c_int64(w), C.c_int64(ldb)))

    def adopt_dsym(self, cohort, U_ptr, w, ldb, ext):
        self._ck(self.lib.sgv_ld_adopt_dsym(self.h, C.c_int(cohort), C.c_void_p(U_ptr), C.c_int64(w), C.c_int64(ldb),
                                            C.c_int64(ext)))

    def dsym_extension(self, w):
        e = C.c_int64()
        self._ck(self.lib.sgv_dsym_extension(self.h, C.c_int64(w), C.byref(e)))
        return e.value

    def adopt_dense(self, cohort, ptr, ld):
        self._ck(self.lib.sgv_ld_adopt_dense(self.h, C.c_int(cohort), C.c_void_p(ptr), C.c_int64(ld)))

    def adopt_blockdiag(self, cohort, ptr, starts, offs, lds):
        st = np.ascontiguousarray(starts, dtype=np.int64)
        of = np.ascontiguousarray(offs, dtype=np.int64)
        ld = np.ascontiguousarray(lds, dtype=np.int32)
        nb = len(of)
        assert len(st) == nb + 1 and len(ld) == nb
        self._ck(self.lib.sgv_ld_adopt_blockdiag(self.h, C.c_int(cohort), C.c_void_p(ptr), C.c_int(nb),
                                                 st.ctypes.data_as(C.POINTER(C.c_int64)), of.ctypes.data_as(C.POINTER(C.c_int64)),
                                                 ld.ctypes.data_as(C.POINTER(C.c_int32))))

    def build_banded(self, cohort, G, N, w, s=0.0, taper=True, y=None, g0=0, device_ptr=None, nmark=None, ldg=None):
        """Banded LD (and XTy) from int8 genotypes, marker-major.  G: (nmark, ldg) int8 host array with ldg % 16 == 0 and
        zero padding beyond N - or device_ptr/nmark/ldg for a buffer already in HBM."""
        out = None
        yp = None
        if y is not None:
            y = _f64(y).ravel()
            assert y.shape[0] == N
            out = np.empty(self.M, dtype=np.float64)
            yp = _dp(y)
        if device_ptr is None:
            G = np.ascontiguousarray(G, dtype=np.int8)
            nmark, ldg = G.shape
            ptr, on_dev = G.ctypes.data_as(C.c_void_p), 0
        else:
            ptr, on_dev = C.c_void_p(int(device_ptr)), 1
        self._ck(self.lib.sgv_ld_build_banded(self.h, C.c_int(cohort), ptr, C.c_int(on_dev), C.c_int64(g0), C.c_int64(nmark),
                                              C.c_int64(N), C.c_int64(ldg), C.c_int64(w), C.c_double(s), C.c_int(int(bool(taper))),
                                              yp, _dp(out) if out is not None else None))
        return out

    def band_shape(self, cohort):
        w, ldb, ext = C.c_int64(), C.c_int64(), C.c_int64()
        self._ck(self.lib.sgv_ld_copy_band(self.h, C.c_int(cohort), None, C.c_int64(0), C.byref(w), C.byref(ldb), C.byref(ext)))
        return w.value, ldb.value, ext.value

    def copy_band(self, cohort, dst_ptr, nfloats):
        self._ck(self.lib.sgv_ld_copy_band(self.h, C.c_int(cohort), C.c_void_p(int(dst_ptr)), C.c_int64(nfloats), None, None, None))

    def ld_info(self, cohort):
        layout = C.c_int()
        nnz = C.c_int64()
        bw = C.c_int64()
        nb = C.c_int64()
        bpp = C.c_double()
        self._ck(self.lib.sgv_ld_info(self.h, C.c_int(cohort), C.byref(layout), C.byref(nnz), C.byref(bw),
                                      C.byref(nb), C.byref(bpp)))
        return dict(layout=LAYOUT_NAMES[layout.value], nnz_stored=nnz.value, bandwidth=bw.value,
                    nblocks=nb.value, bytes_per_pass=bpp.value)

    def set_xty(self, cohort, r):
        r = _f64(r).ravel()
        assert r.shape[0] == self.M
        self._ck(self.lib.sgv_set_xty(self.h, C.c_int(cohort), _dp(r)))

    def reset_state(self):
        self._ck(self.lib.sgv_reset_state(self.h))

    def get_vec(self, cohort, which):
        out = np.empty(self.M, dtype=np.float64)
        self._ck(self.lib.sgv_get_vec(self.h, C.c_int(cohort), C.c_int(which), _dp(out)))
        return out

    def r1_block(self):
        """(device pointer, stride in doubles) of the K x rows block of r1 vectors (see sgv_r1_block)."""
        p, st = C.c_void_p(), C.c_int64()
        self._ck(self.lib.sgv_r1_block(self.h, C.byref(p), C.byref(st)))
        return int(p.value), int(st.value)

    def set_vec(self, cohort, which, v):
        v = _f64(v).ravel()
        assert v.shape[0] == self.M
        self._ck(self.lib.sgv_set_vec(self.h, C.c_int(cohort), C.c_int(which), _dp(v)))

    def pinned_array(self, n):
        p = C.c_void_p()
        self._ck(self.lib.sgv_pinned_alloc(self.h, C.c_int64(n * 8), C.byref(p)))
        self._pinned.append(p.value)
        buf = (C.c_double * n).from_address(p.value)
        return np.frombuffer(buf, dtype=np.float64)

    def get_vec_async(self, cohort, which, scale, pinned):
        self._ck(self.lib.sgv_get_vec_async(self.h, C.c_int(cohort), C.c_int(which), C.c_double(scale), _dp(pinned)))

    def wait_copies(self):
        self._ck(self.lib.sgv_wait_copies(self.h))

    # -- prior ---------------------------------------------------------------------------------
    def set_prior(self, lam, omegas, sigmas):
        om, sg = _f64(omegas), _f64(sigmas)
        self._ck(self.lib.sgv_set_prior(self.h, C.c_int(len(om) + 1), C.c_double(lam), _dp(om), _dp(sg)))

    def set_weights(self, a):
        a = _f64(a)
        assert a.shape[0] == self.K
        self._ck(self.lib.sgv_set_weights(self.h, _dp(a)))

    # -- per-iteration steps -------------------------------------------------------------------
    def denoise(self, gam1s, rho, damp):
        g = _f64(gam1s)
        out = C.c_double()
        self._ck(self.lib.sgv_denoise(self.h, _dp(g), C.c_double(rho), C.c_int(int(damp)), C.byref(out)))
        return out.value

    def prior_em(self, gam1s, maxit, tol, Lm1):
        g = _f64(gam1s)
        lam = C.c_double()
        om = np.zeros(max(Lm1, 1))
        steps = C.c_int()
        rel = C.c_double()
        self._ck(self.lib.sgv_prior_em(self.h, _dp(g), C.c_int(maxit), C.c_double(tol), C.byref(lam), _dp(om),
                                       C.byref(steps), C.byref(rel)))
        return lam.value, om[:Lm1].copy(), steps.value, rel.value

    def lagrangian(self, gam1s, x, omega0, sigma2):
        g, x, o, s2 = _f64(gam1s), _f64(x), _f64(omega0), _f64(sigma2)
        y = np.zeros(x.shape[0])
        self._ck(self.lib.sgv_lagrangian(self.h, _dp(g), _dp(x), _dp(o), _dp(s2), _dp(y)))
        return y

    def lmmse(self, cohort, gamw, gam2, alpha1, rho, cg_maxit, lmmse_damp, learn_gamw, x0_zero, probe):
        pin = LmmseIn(gamw, gam2, alpha1, rho, int(cg_maxit), int(bool(lmmse_damp)), int(bool(learn_gamw)),
                      int(bool(x0_zero)))
        out = LmmseOut()
        probe = np.ascontiguousarray(probe, dtype=np.int8)
        assert probe.shape[0] == self.M
        self._ck(self.lib.sgv_lmmse(self.h, C.c_int(cohort), C.byref(pin), probe.ctypes.data_as(C.POINTER(C.c_int8)),
                                    C.byref(out)))
        return out

    def probe_pair(self, cohort, gamw, gam2, cg_maxit, probe_a, probe_b=None):
        """Two further Hutchinson probes through the 2-RHS solver (see sgv_probe_pair)."""
        out = ProbeOut()
        pa = np.ascontiguousarray(probe_a, dtype=np.int8)
        assert pa.shape[0] == self.M
        pb = None
        if probe_b is not None:
            pb = np.ascontiguousarray(probe_b, dtype=np.int8)
            assert pb.shape[0] == self.M
        self._ck(self.lib.sgv_probe_pair(self.h, C.c_int(cohort), C.c_double(gamw), C.c_double(gam2), C.c_int(int(cg_maxit)),
                                         pa.ctypes.data_as(C.POINTER(C.c_int8)),
                                         pb.ctypes.data_as(C.POINTER(C.c_int8)) if pb is not None else None, C.byref(out)))
        return out

    def update_r1(self, cohort, alpha2):
        self._ck(self.lib.sgv_update_r1(self.h, C.c_int(cohort), C.c_double(alpha2)))

    def metrics(self, x0=None):
        d = np.zeros(4)
        p = _dp(_f64(x0).ravel()) if x0 is not None else None
        self._ck(self.lib.sgv_metrics(self.h, p, _dp(d)))
        return d

    # -- fused VAMP iteration (device-resident scalar chain) -------------------------------------
    def iteration_supported(self):
        return bool(self.lib.sgv_iteration_supported(self.h))

    def vamp_begin(self, gam1, gamw, N):
        g1, gw, n = _f64(gam1), _f64(gamw), _f64(N)
        assert g1.shape[0] == gw.shape[0] == n.shape[0] == self.K
        self._ck(self.lib.sgv_vamp_begin(self.h, _dp(g1), _dp(gw), _dp(n)))

    def vamp_set_alphas(self, alpha1, alpha2):
        a1, a2 = _f64(alpha1), _f64(alpha2)
        self._ck(self.lib.sgv_vamp_set_alphas(self.h, _dp(a1), _dp(a2)))

    def set_truth(self, x0):
        x0 = _f64(x0).ravel()
        assert x0.shape[0] == self.M
        self._ck(self.lib.sgv_set_truth(self.h, _dp(x0)))

    def iteration_probe_buffer(self, slot):
        """The pinned (K, local rows) int8 staging buffer of a slot: write the iteration's probes into it."""
        p = C.POINTER(C.c_int8)()
        self._ck(self.lib.sgv_iteration_probe_buffer(self.h, C.c_int(slot), C.byref(p)))
        n = self.K * self.M
        return np.ctypeslib.as_array(p, shape=(n,)).reshape(self.K, self.M)

    def iteration_enqueue(self, it, update_prior, em_maxit, em_tol, rho, cg_maxit, lmmse_damp, learn_gamw, want_metrics,
                          xhat_pinned, r1_pinned, slot):
        pin = IterIn(int(it), int(bool(update_prior)), int(em_maxit), float(em_tol), float(rho), int(cg_maxit),
                     int(bool(lmmse_damp)), int(bool(learn_gamw)), int(bool(want_metrics)))
        xp = _dp(xhat_pinned) if xhat_pinned is not None else None
        rp = None
        if r1_pinned is not None:
            arr = (C.POINTER(C.c_double) * self.K)()
            for k in range(self.K):
                arr[k] = _dp(r1_pinned[k]) if r1_pinned[k] is not None else None
            rp = arr
        self._ck(self.lib.sgv_iteration_enqueue(self.h, C.byref(pin), xp, rp, C.c_int(slot)))

    def iteration_wait(self, slot):
        out = IterOut()
        self._ck(self.lib.sgv_iteration_wait(self.h, C.c_int(slot), C.byref(out)))
        return out

    # -- hooks ---------------------------------------------------------------------------------
    def spmm(self, cohort, X, alpha=1.0, beta=0.0):
        X = _f64(X)
        nrhs = 1 if X.ndim == 1 else X.shape[1]
        Xf = np.ascontiguousarray(X.reshape(self.M, nrhs).T)     # column-major M x nrhs
        Y = np.zeros_like(Xf)
        self._ck(self.lib.sgv_spmm(self.h, C.c_int(cohort), _dp(Xf), _dp(Y), C.c_int(nrhs), C.c_double(alpha),
                                   C.c_double(beta)))
        return Y.T.reshape(X.shape).copy()

    def spmm_stage(self, X):
        X = _f64(X)
        nrhs = 1 if X.ndim == 1 else X.shape[1]
        Xf = np.ascontiguousarray(X.reshape(self.M, nrhs).T)
        self._ck(self.lib.sgv_spmm_stage(self.h, _dp(Xf), C.c_int(nrhs)))
        return nrhs

    def spmm_run(self, cohort, nrhs, alpha=1.0, beta=0.0):
        Y = np.zeros((nrhs, self.M))
        self._ck(self.lib.sgv_spmm_run(self.h, C.c_int(cohort), _dp(Y), C.c_int(nrhs), C.c_double(alpha), C.c_double(beta)))
        return Y.T.copy() if nrhs == 2 else Y[0].copy()

    def spmm_bench(self, cohort, reps):
        ms = C.c_float()
        self._ck(self.lib.sgv_spmm_bench(self.h, C.c_int(cohort), C.c_int(reps), C.byref(ms)))
        return ms.value

    def profile(self, enable):
        self._ck(self.lib.sgv_profile(self.h, C.c_int(int(enable))))

    def profile_read(self):
        ms = C.c_double()
        n = C.c_int64()
        self._ck(self.lib.sgv_profile_read(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def launch_count(self):
        return int(self.lib.sgv_launch_count(self.h))
