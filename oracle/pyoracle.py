"""ctypes front ends for the two CPU oracles (TEST INFRASTRUCTURE ONLY).

* ``RefOracle``  -- oracle/_ref/libemu_ref.so: the reference's own C sources compiled unmodified
  against oracle/gsl_shim (built by oracle/Makefile where /root/reference exists; the prebuilt .so
  travels to the GPU box).
* ``PortOracle`` -- oracle/libemu_oracle.so: our plain-C restatement (oracle/emu_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module; nothing
under madaiemulator_b200/ does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = ctypes.POINTER(ctypes.c_double)
_vp = ctypes.c_void_p
_ci = ctypes.c_int

POWEREXP, MATERN32, MATERN52 = 1, 2, 3
DET_LOGSUM, DET_PRODUCT = 0, 1


def _P(a):
    return a.ctypes.data_as(_dp)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def build(ref=True):
    """Compile the oracles (the reference build only where /root/reference is present)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "libemu_oracle.so"])
    if ref and os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-C", _HERE, "_ref/libemu_ref.so"])


def ref_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libemu_ref.so"))


class PortOracle:
    """Plain-C restatement (emu_oracle.c)."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            path = os.path.join(_HERE, "libemu_oracle.so")
            if not os.path.exists(path):
                build(ref=False)
            L = ctypes.CDLL(path)
            L.emuo_cov_pair.restype = ctypes.c_double
            L.emuo_cov_pair.argtypes = [_ci, _dp, _dp, _dp, _ci]
            L.emuo_cov_matrix.argtypes = [_ci, _dp, _ci, _ci, _dp, _dp]
            L.emuo_deriv_matrix.argtypes = [_ci, _dp, _ci, _ci, ctypes.c_double, _ci, _dp]
            L.emuo_k_vector.argtypes = [_ci, _dp, _ci, _ci, _dp, _dp, _dp]
            L.emuo_h_matrix.argtypes = [_ci, _dp, _ci, _ci, _dp]
            L.emuo_cholesky.argtypes = [_dp, _ci]
            L.emuo_cholesky_invert.argtypes = [_dp, _ci]
            L.emuo_estimate_beta.argtypes = [_dp, _dp, _dp, _ci, _ci, _dp]
            L.emuo_ranges.argtypes = [_ci, _dp, _ci, _ci, _dp]
            L.emuo_sample_scales.argtypes = [_dp, _ci, _ci, _dp]
            L.emuo_loglik_grad.argtypes = [_ci, _ci, _dp, _ci, _ci, _dp, _dp, _ci, _dp, _dp, _dp, _dp, _dp]
            L.emuo_emulator_create.restype = _vp
            L.emuo_emulator_create.argtypes = [_ci, _ci, _dp, _ci, _ci, _dp, _dp]
            L.emuo_emulator_free.argtypes = [_vp]
            L.emuo_emulate.argtypes = [_vp, _dp, _ci, _dp, _dp]
            L.emuo_emulator_beta.argtypes = [_vp, _dp]
            L.emuo_backproject.argtypes = [_ci, _ci, _dp, _dp, _dp, _dp, _dp, _dp, _dp]
            cls._lib = L
        return cls._lib

    def __init__(self, X, y, kernel=POWEREXP, order=0):
        self.X = _c(X)
        self.y = _c(y)
        self.n, self.d = self.X.shape
        self.kernel, self.order = kernel, order
        self.nthetas = 3 if kernel in (MATERN32, MATERN52) else self.d + 2
        self.p = 1 + (order if 0 <= order <= 3 else 0) * self.d
        self.L = self.lib()

    def cov_matrix(self, thetas):
        C = np.empty((self.n, self.n))
        self.L.emuo_cov_matrix(self.kernel, _P(self.X), self.n, self.d, _P(_c(thetas)), _P(C))
        return C

    def cov_pair(self, xa, xb, thetas):
        return self.L.emuo_cov_pair(self.kernel, _P(_c(xa)), _P(_c(xb)), _P(_c(thetas)), self.d)

    def deriv_matrix(self, theta_length, index):
        D = np.empty((self.n, self.n))
        self.L.emuo_deriv_matrix(self.kernel, _P(self.X), self.n, self.d, float(theta_length), index, _P(D))
        return D

    def h_matrix(self):
        H = np.empty((self.n, self.p))
        self.L.emuo_h_matrix(self.order, _P(self.X), self.n, self.d, _P(H))
        return H

    def ranges(self):
        r = np.empty((self.nthetas, 2))
        self.L.emuo_ranges(self.kernel, _P(self.X), self.n, self.d, _P(r))
        return r

    def loglik_grad(self, theta_less_amp, det_mode=DET_LOGSUM, want_grad=True):
        th = _c(theta_less_amp)
        negL = ctypes.c_double()
        s2 = ctypes.c_double()
        ld = ctypes.c_double()
        g = np.zeros(self.nthetas - 1)
        beta = np.zeros(self.p)
        rc = self.L.emuo_loglik_grad(self.kernel, self.order, _P(self.X), self.n, self.d, _P(self.y), _P(th),
                                     det_mode, ctypes.byref(negL), _P(g) if want_grad else None,
                                     ctypes.byref(s2), ctypes.byref(ld), _P(beta))
        return dict(status=rc, negL=negL.value, grad=g if want_grad else None, sigma2=s2.value,
                    logdet=ld.value, beta=beta)

    def emulator(self, thetas):
        return _PortEmulator(self, thetas)


class _PortEmulator:
    def __init__(self, o, thetas):
        self.o = o
        self.h = o.L.emuo_emulator_create(o.kernel, o.order, _P(o.X), o.n, o.d, _P(o.y), _P(_c(thetas)))
        if not self.h:
            raise ValueError("covariance or regression matrix not positive definite")

    def emulate(self, pts):
        pts = _c(pts).reshape(-1, self.o.d)
        m = pts.shape[0]
        mean, var = np.empty(m), np.empty(m)
        self.o.L.emuo_emulate(self.h, _P(pts), m, _P(mean), _P(var))
        return mean, var

    def beta(self):
        b = np.empty(self.o.p)
        self.o.L.emuo_emulator_beta(self.h, _P(b))
        return b

    def __del__(self):
        if getattr(self, "h", None):
            self.o.L.emuo_emulator_free(self.h)
            self.h = None


def backproject(training_mean, evecs, evals, mean_pca, var_pca):
    L = PortOracle.lib()
    evecs = _c(evecs)
    nt, nr = evecs.shape
    mo, vo = np.empty(nt), np.empty(nt)
    L.emuo_backproject(nt, nr, _P(_c(training_mean)), _P(evecs), _P(_c(evals)), _P(_c(mean_pca)),
                       _P(_c(var_pca)), _P(mo), _P(vo))
    return mo, vo


def dropin_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libemu_dropin.so"))


class RefOracle:
    """The reference's own sources (oracle/_ref/libemu_ref.so)."""

    _lib = None
    _libname = "libemu_ref.so"

    @classmethod
    def lib(cls):
        if cls.__dict__.get("_lib") is None:
            path = os.path.join(_HERE, "_ref", cls._libname)
            if not os.path.exists(path):
                build(ref=True)
            L = ctypes.CDLL(path)
            L.ref_model_create.restype = _vp
            L.ref_model_create.argtypes = [_dp, _ci, _ci, _dp, _ci, _ci]
            L.ref_model_free.argtypes = [_vp]
            L.ref_model_nthetas.argtypes = [_vp]
            L.ref_model_nregression_fns.argtypes = [_vp]
            L.ref_model_ranges.argtypes = [_vp, _dp]
            L.ref_model_sample_scales.argtypes = [_vp, _dp]
            if hasattr(L, "ref_model_ranges_ex"):
                L.ref_model_ranges_ex.argtypes = [_vp, _ci, _ci, ctypes.c_double, _dp]
                L.ref_model_ranges_ex.restype = None
            L.ref_cov_matrix.argtypes = [_vp, _dp, _dp]
            if hasattr(L, "ref_k_vector"):
                L.ref_k_vector.argtypes = [_vp, _dp, _dp, _dp]
                L.ref_k_vector.restype = None
                L.ref_chol_inverse.argtypes = [_vp, _dp, _dp, _dp]
                L.ref_chol_inverse.restype = None
            L.ref_cov_pair.restype = ctypes.c_double
            L.ref_cov_pair.argtypes = [_vp, _dp, _dp, _dp]
            L.ref_deriv_matrix.argtypes = [_vp, ctypes.c_double, _ci, _dp]
            L.ref_h_matrix.argtypes = [_vp, _dp]
            L.ref_eval.restype = ctypes.c_double
            L.ref_eval.argtypes = [_vp, _dp]
            L.ref_grad.argtypes = [_vp, _dp, _dp]
            L.ref_eval_grad.argtypes = [_vp, _dp, _dp, _dp]
            L.ref_sigma_full.restype = ctypes.c_double
            L.ref_sigma_full.argtypes = [_vp, _dp]
            L.ref_eval_logsum.argtypes = [_vp, _dp, _dp, _dp, _dp, _dp]
            L.ref_cinverse.argtypes = [_vp, _dp, _dp]
            L.ref_emulator_create.restype = _vp
            L.ref_emulator_create.argtypes = [_vp, _dp]
            L.ref_emulator_free.argtypes = [_vp]
            L.ref_emulate.argtypes = [_vp, _dp, _ci, _dp, _dp]
            if hasattr(L, "ref_emulate_at_point_list"):
                L.ref_emulate_at_point_list.argtypes = [_vp, _dp, _dp, _ci, _ci, _dp, _dp]
                L.ref_emulate_at_point_list.restype = None
            if hasattr(L, "ref_emulate_model_results"):
                L.ref_emulate_model_results.argtypes = [_vp, _dp, _dp, _ci, _dp, _dp]
                L.ref_emulate_model_results.restype = None
            L.ref_emulator_beta.argtypes = [_vp, _dp]
            L.ref_max_with_multimin.restype = ctypes.c_double
            L.ref_max_with_multimin.argtypes = [_vp, _ci, ctypes.c_ulong, _dp]
            L.ref_random_inits.argtypes = [_vp, ctypes.c_ulong, _ci, _dp]
            L.ref_time_eval_grad.restype = ctypes.c_double
            L.ref_time_eval_grad.argtypes = [ctypes.POINTER(_vp), _ci, _dp, _ci]
            L.ref_time_emulate.restype = ctypes.c_double
            L.ref_time_emulate.argtypes = [_vp, _dp, _ci, _ci, _dp, _dp]
            L.ref_ncpus.restype = _ci
            cls._lib = L
        return cls.__dict__["_lib"]

    def __init__(self, X, y, kernel=POWEREXP, order=0):
        self.L = self.lib()
        self.X = _c(X)
        self.y = _c(y)
        self.n, self.d = self.X.shape
        self.kernel, self.order = kernel, order
        self.h = self.L.ref_model_create(_P(self.X), self.n, self.d, _P(self.y), kernel, order)
        self.nthetas = self.L.ref_model_nthetas(self.h)
        self.p = self.L.ref_model_nregression_fns(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_model_free(self.h)
            self.h = None

    def cov_matrix(self, thetas):
        C = np.empty((self.n, self.n))
        self.L.ref_cov_matrix(self.h, _P(_c(thetas)), _P(C))
        return C

    def cov_pair(self, xa, xb, thetas):
        return self.L.ref_cov_pair(self.h, _P(_c(xa)), _P(_c(xb)), _P(_c(thetas)))

    def k_vector(self, thetas, xnew):
        """makeKVector_fnptr (emulator.c:578)"""
        k = np.empty(self.n)
        self.L.ref_k_vector(self.h, _P(_c(thetas)), _P(_c(xnew)), _P(k))
        return k

    def chol_inverse(self, A):
        """chol_inverse_cov_matrix (emulate-fns.c:275): (inverse, determinant) of an n x n matrix"""
        A = _c(A)
        out = np.empty_like(A)
        det = ctypes.c_double()
        self.L.ref_chol_inverse(self.h, _P(A), _P(out), ctypes.byref(det))
        return out, det.value

    def deriv_matrix(self, theta_length, index):
        D = np.empty((self.n, self.n))
        self.L.ref_deriv_matrix(self.h, float(theta_length), index, _P(D))
        return D

    def h_matrix(self):
        H = np.empty((self.n, self.p))
        self.L.ref_h_matrix(self.h, _P(H))
        return H

    def ranges(self):
        r = np.empty((self.nthetas, 2))
        self.L.ref_model_ranges(self.h, _P(r))
        return r

    def ranges_ex(self, use_data_scales=True, fixed_nugget=None):
        r = np.empty((self.nthetas, 2))
        self.L.ref_model_ranges_ex(self.h, 1 if use_data_scales else 0, 0 if fixed_nugget is None else 1,
                                   0.0 if fixed_nugget is None else float(fixed_nugget), _P(r))
        return r

    def eval(self, theta_less_amp):
        """Literal evalFnMulti: -L with the running-product determinant."""
        return self.L.ref_eval(self.h, _P(_c(theta_less_amp)))

    def grad(self, theta_less_amp):
        g = np.empty(self.nthetas - 1)
        self.L.ref_grad(self.h, _P(_c(theta_less_amp)), _P(g))
        return g

    def eval_logsum(self, theta_less_amp):
        negL, ld, s2 = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        beta = np.empty(self.p)
        rc = self.L.ref_eval_logsum(self.h, _P(_c(theta_less_amp)), ctypes.byref(negL), ctypes.byref(ld),
                                    ctypes.byref(s2), _P(beta))
        return dict(status=rc, negL=negL.value, logdet=ld.value, sigma2=s2.value, beta=beta)

    def cinverse(self, theta_less_amp):
        C = np.empty((self.n, self.n))
        rc = self.L.ref_cinverse(self.h, _P(_c(theta_less_amp)), _P(C))
        return rc, C

    def sigma_full(self, theta_less_amp):
        return self.L.ref_sigma_full(self.h, _P(_c(theta_less_amp)))

    def emulator(self, thetas):
        return _RefEmulator(self, thetas)

    def emulate_at_point_list(self, thetas, pts, single=False):
        """emulateAtPointList (emulate-fns.c:73); single=True: emulateAtPoint (:138) point by point"""
        pts = _c(pts).reshape(-1, self.d)
        m = pts.shape[0]
        mean, var = np.empty(m), np.empty(m)
        self.L.ref_emulate_at_point_list(self.h, _P(_c(thetas)), _P(pts), m, 1 if single else 0, _P(mean), _P(var))
        return mean, var

    def emulate_model_results(self, thetas, pts):
        """emulate_model_results (emulate-fns.c:13) over the rows of pts"""
        pts = _c(pts).reshape(-1, self.d)
        m = pts.shape[0]
        mean, var = np.empty(m), np.empty(m)
        self.L.ref_emulate_model_results(self.h, _P(_c(thetas)), _P(pts), m, _P(mean), _P(var))
        return mean, var

    def max_with_multimin(self, max_tries, seed):
        th = np.empty(self.nthetas)
        best = self.L.ref_max_with_multimin(self.h, max_tries, seed, _P(th))
        return best, th

    def random_inits(self, seed, count):
        out = np.empty((count, self.nthetas))
        self.L.ref_random_inits(self.h, seed, count, _P(out))
        return out


class _RefEmulator:
    def __init__(self, o, thetas):
        self.o = o
        self.h = o.L.ref_emulator_create(o.h, _P(_c(thetas)))

    def emulate(self, pts):
        pts = _c(pts).reshape(-1, self.o.d)
        m = pts.shape[0]
        mean, var = np.empty(m), np.empty(m)
        self.o.L.ref_emulate(self.h, _P(pts), m, _P(mean), _P(var))
        return mean, var

    def beta(self):
        b = np.empty(self.o.p)
        self.o.L.ref_emulator_beta(self.h, _P(b))
        return b

    def __del__(self):
        if getattr(self, "h", None):
            self.o.L.ref_emulator_free(self.h)
            self.h = None


def time_ref_eval_grad(X, y, theta_less_amp, kernel=POWEREXP, order=0, nthreads=1, reps=1):
    """Wall seconds for nthreads*reps evalFnGradMulti calls, one independent model per thread."""
    models = [RefOracle(X, y, kernel, order) for _ in range(nthreads)]
    arr = (_vp * nthreads)(*[m.h for m in models])
    th = _c(theta_less_amp)
    return RefOracle.lib().ref_time_eval_grad(arr, nthreads, _P(th), reps)


def time_ref_emulate(X, y, thetas, pts, kernel=POWEREXP, order=0, nthreads=1):
    o = RefOracle(X, y, kernel, order)
    e = o.emulator(thetas)
    pts = _c(pts)
    m = pts.shape[0]
    mean, var = np.empty(m), np.empty(m)
    t = RefOracle.lib().ref_time_emulate(e.h, _P(pts), m, nthreads, _P(mean), _P(var))
    return t, mean, var


class DropinOracle(RefOracle):
    """The reference's own sources with the hot-path libEmu symbols replaced by integration/libemu_glue.c, i.e. the
    reference running on top of the CUDA engine (oracle/_ref/libemu_dropin.so; needs a B200)."""

    _lib = None
    _libname = "libemu_dropin.so"

    @classmethod
    def reset(cls):
        L = cls.lib()
        L.libemu_glue_reset.restype = None
        L.libemu_glue_reset()
