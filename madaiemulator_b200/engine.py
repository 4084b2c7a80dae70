"""ctypes binding of the C-ABI in include/emu_b200.h (madaiemulator_b200/csrc/libemub.so).

This is the host-side mirror of the reference's libEmu operator interface for the hot path
(evalFnMulti / gradFnMulti / makeCovMatrix / alloc_emulator_struct / emulate_point, see the header for
file:line citations).  There is no CPU fallback: if the CUDA library is missing or no B200 is visible
every call raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libemub.so")

POWEREXP, MATERN32, MATERN52 = 1, 2, 3
MAXD, MAXNCP = 64, 104               # nparams / (1 + nregression_fns) limits of the engine (csrc/emub_kernels.cuh)
RES_STRIDE = 8 + MAXNCP + MAXD + 2   # doubles per point the host-pointer likelihood call reads back
OK, EDOM, EREG, EINVAL, ECUDA, ENOMEM = 0, 1, 2, 3, 4, 5

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)
_vp = ctypes.c_void_p
_ci = ctypes.c_int
_ll = ctypes.c_longlong

FAMILIES = ["cov", "potf2", "gemm_chol", "gemm_trtri", "gemm_lauum", "skinny", "small", "grad", "kcross",
            "gemm_pred", "pred_final"]

# every symbol include/emu_b200.h declares (checked by tests/test_cabi_symbols.py)
SYMBOLS = [
    "emub_ctx_create", "emub_ctx_destroy", "emub_last_error", "emub_version", "emub_ctx_stream",
    "emub_ctx_set_groups", "emub_ctx_use_graphs", "emub_model_create", "emub_model_destroy", "emub_model_nthetas",
    "emub_model_nregression_fns", "emub_model_slots", "emub_model_kernel", "emub_spd_inverse", "emub_model_set_training", "emub_model_set_training_multi",
    "emub_model_ncomponents", "emub_model_set_gradient_mode", "emub_model_gradient_mode", "emub_loglik_grad_batch_comp", "emub_loglik_grad_batch_mixed", "emub_emulator_create_comp", "emub_predict_multi", "emub_predict_multi_few", "emub_cov_matrix",
    "emub_h_matrix", "emub_k_vectors", "emub_loglik_grad_batch", "emub_loglik_grad_batch_dev",
    "emub_ctx_synchronize", "emub_loglik_extras", "emub_emulator_create", "emub_emulator_destroy",
    "emub_emulator_beta", "emub_predict_batch", "emub_predict_few", "emub_predict_batch_dev", "emub_profile_enable",
    "emub_profile_reset", "emub_profile_read", "emub_profile_name", "emub_launch_count", "emub_debug_fetch",
    "emub_debug_cholesky", "emub_debug_exp", "emub_debug_exp_scaled",
]


class EmubError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("emub error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib():
    """Load libemub.so (raises if it has not been built: there is no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("CUDA engine %s is missing: run `make -C madaiemulator_b200/csrc` "
                           "(or __graft_entry__.build()); there is no CPU fallback" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    L.emub_last_error.restype = ctypes.c_char_p
    L.emub_version.restype = ctypes.c_char_p
    L.emub_profile_name.restype = ctypes.c_char_p
    L.emub_profile_name.argtypes = [_ci]
    L.emub_ctx_create.argtypes = [_ci, ctypes.POINTER(_vp)]
    L.emub_ctx_destroy.argtypes = [_vp]
    L.emub_ctx_destroy.restype = None
    L.emub_ctx_stream.argtypes = [_vp]
    L.emub_ctx_stream.restype = _vp
    L.emub_model_set_gradient_mode.argtypes = [_vp, _ci]
    L.emub_model_gradient_mode.argtypes = [_vp]
    L.emub_ctx_set_groups.argtypes = [_vp, _ci]
    L.emub_ctx_use_graphs.argtypes = [_vp, _ci]
    L.emub_ctx_synchronize.argtypes = [_vp]
    L.emub_model_create.argtypes = [_vp, _dp, _ci, _ci, _ci, _dp, _ci, _ci, _ci, ctypes.POINTER(_vp)]
    L.emub_model_destroy.argtypes = [_vp]
    L.emub_model_destroy.restype = None
    L.emub_model_nthetas.argtypes = [_vp]
    L.emub_model_nregression_fns.argtypes = [_vp]
    L.emub_model_slots.argtypes = [_vp]
    L.emub_model_kernel.argtypes = [_vp]
    L.emub_spd_inverse.argtypes = [_vp, _dp, _ci, _dp, _ci, _dp]
    L.emub_model_set_training.argtypes = [_vp, _dp]
    L.emub_model_set_training_multi.argtypes = [_vp, _dp, _ci, _ci]
    L.emub_model_ncomponents.argtypes = [_vp]
    L.emub_loglik_grad_batch_comp.argtypes = [_vp, _dp, _ip, _ci, _ci, _dp, _dp, _dp, _ip]
    L.emub_loglik_grad_batch_mixed.argtypes = [_vp, _dp, _ip, _ip, _ci, _dp, _dp, _dp, _ip]
    L.emub_emulator_create_comp.argtypes = [_vp, _ci, _dp, ctypes.POINTER(_vp)]
    L.emub_predict_multi.argtypes = [ctypes.POINTER(_vp), _ci, _dp, _ci, _ci, _ci, _dp, _dp, _dp, _dp, _dp]
    L.emub_predict_multi_few.argtypes = [ctypes.POINTER(_vp), _ci, _dp, _ci, _ci, _ci, _dp, _dp, _dp, _dp, _dp]
    L.emub_cov_matrix.argtypes = [_vp, _dp, _dp, _ci]
    L.emub_h_matrix.argtypes = [_vp, _dp, _ci]
    L.emub_k_vectors.argtypes = [_vp, _dp, _dp, _ci, _ci, _dp, _ci]
    L.emub_loglik_grad_batch.argtypes = [_vp, _dp, _ci, _ci, _dp, _dp, _dp, _ip]
    L.emub_loglik_grad_batch_dev.argtypes = [_vp, _vp, _ci, _ci, _vp]
    L.emub_loglik_extras.argtypes = [_vp, _ci, _dp, _dp]
    L.emub_emulator_create.argtypes = [_vp, _dp, ctypes.POINTER(_vp)]
    L.emub_emulator_destroy.argtypes = [_vp]
    L.emub_emulator_destroy.restype = None
    L.emub_emulator_beta.argtypes = [_vp, _dp]
    L.emub_predict_batch.argtypes = [_vp, _dp, _ci, _ci, _dp, _dp]
    L.emub_predict_few.argtypes = [_vp, _dp, _ci, _ci, _dp, _dp]
    L.emub_predict_batch_dev.argtypes = [_vp, _vp, _ci, _vp, _vp]
    L.emub_profile_enable.argtypes = [_vp, _ci]
    L.emub_profile_reset.argtypes = [_vp]
    L.emub_profile_read.argtypes = [_vp, _ci, _dp, ctypes.POINTER(_ll), _dp]
    L.emub_launch_count.argtypes = [_vp]
    L.emub_launch_count.restype = _ll
    L.emub_debug_fetch.argtypes = [_vp, _ci, _ci, _dp, _ci]
    L.emub_debug_cholesky.argtypes = [_vp, _dp, _dp, _ci, _dp]
    L.emub_debug_exp.argtypes = [_vp, _dp, _ci, _dp]
    L.emub_debug_exp_scaled.argtypes = [_vp, _dp, _ci, _dp]
    _lib = L
    return L


def _P(a):
    return a.ctypes.data_as(_dp)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _check(rc):
    if rc != OK:
        raise EmubError(rc, lib().emub_last_error().decode())


class Context:
    """One GPU (emub_ctx)."""

    def __init__(self, device=0, groups=None):
        self.L = lib()
        h = _vp()
        _check(self.L.emub_ctx_create(device, ctypes.byref(h)))
        self.h = h
        self.device = device
        if groups is not None:
            self.set_groups(groups)

    def set_groups(self, g):
        _check(self.L.emub_ctx_set_groups(self.h, g))

    def use_graphs(self, on):
        _check(self.L.emub_ctx_use_graphs(self.h, 1 if on else 0))

    def stream(self):
        return self.L.emub_ctx_stream(self.h)

    def synchronize(self):
        _check(self.L.emub_ctx_synchronize(self.h))

    def debug_exp(self, x):
        x = _c(x).ravel()
        out = np.empty_like(x)
        _check(self.L.emub_debug_exp(self.h, _P(x), x.size, _P(out)))
        return out

    def debug_exp_scaled(self, x):
        x = _c(x).ravel()
        out = np.empty_like(x)
        _check(self.L.emub_debug_exp_scaled(self.h, _P(x), x.size, _P(out)))
        return out

    def launch_count(self):
        return int(self.L.emub_launch_count(self.h))

    def profile(self, on):
        _check(self.L.emub_profile_enable(self.h, 1 if on else 0))
        _check(self.L.emub_profile_reset(self.h))

    def profile_read(self):
        out = {}
        for i, name in enumerate(FAMILIES):
            ms, n, w = ctypes.c_double(), _ll(), ctypes.c_double()
            _check(self.L.emub_profile_read(self.h, i, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(w)))
            out[name] = dict(ms=ms.value, launches=n.value, work=w.value)
        return out

    def close(self):
        if getattr(self, "h", None):
            self.L.emub_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Model:
    """Design + training vector + kernel + regression order on the device (emub_model)."""

    def __init__(self, ctx, X, y, kernel=POWEREXP, order=0, max_slots=0):
        self.ctx = ctx
        self.L = ctx.L
        self.X = _c(X)
        self.y = _c(y)
        self.n, self.d = self.X.shape
        h = _vp()
        _check(self.L.emub_model_create(ctx.h, _P(self.X), self.d, self.n, self.d, _P(self.y), kernel, order,
                                        max_slots, ctypes.byref(h)))
        self.h = h
        self.kernel, self.order = kernel, order
        self.nthetas = self.L.emub_model_nthetas(h)
        self.p = self.L.emub_model_nregression_fns(h)
        self.slots = self.L.emub_model_slots(h)

    def set_gradient_mode(self, exact):
        """False: the reference's literal gradient formula (default); True: the true gradient of -L (deviation D-4)."""
        _check(self.L.emub_model_set_gradient_mode(self.h, 1 if exact else 0))

    def set_training(self, y):
        self.y = _c(y)
        _check(self.L.emub_model_set_training(self.h, _P(self.y)))

    def set_training_multi(self, Y):
        """Y: n x ncomp (e.g. the PCA z-matrix): several training vectors on the same design."""
        Y = _c(Y).reshape(self.n, -1)
        _check(self.L.emub_model_set_training_multi(self.h, _P(Y), Y.shape[1], Y.shape[1]))
        self.Y = Y
        self.ncomp = Y.shape[1]

    def cov_matrix(self, thetas):
        C = np.empty((self.n, self.n))
        _check(self.L.emub_cov_matrix(self.h, _P(_c(thetas)), _P(C), self.n))
        return C

    def h_matrix(self):
        H = np.empty((self.n, self.p))
        _check(self.L.emub_h_matrix(self.h, _P(H), self.p))
        return H

    def k_vectors(self, thetas, pts):
        pts = _c(pts).reshape(-1, self.d)
        K = np.empty((self.n, pts.shape[0]))
        _check(self.L.emub_k_vectors(self.h, _P(_c(thetas)), _P(pts), self.d, pts.shape[0], _P(K), pts.shape[0]))
        return K

    def loglik_grad_batch(self, thetas, want_grad=True, comp=None):
        """thetas: B x (nthetas-1); comp: optional training-vector index per point.
        Returns dict(negL[B], grad[B, nthetas-1], sigma2[B], status[B])."""
        th = _c(thetas).reshape(-1, self.nthetas - 1)
        B = th.shape[0]
        negL, s2 = np.empty(B), np.empty(B)
        grad = np.zeros((B, self.nthetas - 1))
        st = np.zeros(B, dtype=np.int32)
        cp = None
        if comp is not None:
            comp = np.ascontiguousarray(comp, dtype=np.int32)
            cp = comp.ctypes.data_as(_ip)
        if isinstance(want_grad, (list, tuple, np.ndarray)):  # a gradient request per point (emub_loglik_grad_batch_mixed)
            wg = np.ascontiguousarray(want_grad, dtype=np.int32)
            assert wg.shape == (B,)
            _check(self.L.emub_loglik_grad_batch_mixed(self.h, _P(th), cp, wg.ctypes.data_as(_ip), B, _P(negL), _P(grad), _P(s2),
                                                       st.ctypes.data_as(_ip)))
            return dict(negL=negL, grad=grad, sigma2=s2, status=st)
        _check(self.L.emub_loglik_grad_batch_comp(self.h, _P(th), cp, B, 1 if want_grad else 0, _P(negL), _P(grad), _P(s2),
                                                  st.ctypes.data_as(_ip)))
        return dict(negL=negL, grad=grad, sigma2=s2, status=st)

    def loglik_grad(self, theta_less_amp, want_grad=True):
        r = self.loglik_grad_batch(np.asarray(theta_less_amp)[None, :], want_grad)
        ld, beta = ctypes.c_double(), np.zeros(self.p)
        _check(self.L.emub_loglik_extras(self.h, 0, ctypes.byref(ld), _P(beta)))
        return dict(status=int(r["status"][0]), negL=float(r["negL"][0]), grad=r["grad"][0], sigma2=float(r["sigma2"][0]),
                    logdet=ld.value, beta=beta)

    def loglik_grad_batch_dev(self, d_thetas_ptr, B, want_grad, d_out_ptr):
        _check(self.L.emub_loglik_grad_batch_dev(self.h, d_thetas_ptr, B, 1 if want_grad else 0, d_out_ptr))

    def spd_inverse(self, A):
        """chol_inverse_cov_matrix (emulate-fns.c:275) for a caller-owned n x n matrix: (A^-1, log det A)."""
        A = _c(A)
        out = np.empty_like(A)
        ld = ctypes.c_double()
        _check(self.L.emub_spd_inverse(self.h, _P(A), self.n, _P(out), self.n, ctypes.byref(ld)))
        return out, ld.value

    def debug_fetch(self, slot, which):
        out = np.empty((self.n, self.n))
        _check(self.L.emub_debug_fetch(self.h, slot, which, _P(out), self.n))
        return out

    def debug_cholesky(self, theta_less_amp):
        Lm = np.zeros((self.n, self.n))
        ld = ctypes.c_double()
        rc = self.L.emub_debug_cholesky(self.h, _P(_c(theta_less_amp)), _P(Lm), self.n, ctypes.byref(ld))
        if rc not in (OK, EDOM):
            _check(rc)
        return rc, Lm, ld.value

    def emulator(self, thetas, comp=0):
        return Emulator(self, thetas, comp)

    def close(self):
        if getattr(self, "h", None):
            self.L.emub_model_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Emulator:
    """Cached factor for prediction (emub_emulator; the reference's emulator_struct)."""

    def __init__(self, model, thetas, comp=0):
        self.model = model
        self.L = model.L
        h = _vp()
        _check(self.L.emub_emulator_create_comp(model.h, comp, _P(_c(thetas)), ctypes.byref(h)))
        self.h = h

    def emulate(self, pts):
        pts = _c(pts).reshape(-1, self.model.d)
        m = pts.shape[0]
        mean, var = np.empty(m), np.empty(m)
        _check(self.L.emub_predict_batch(self.h, _P(pts), self.model.d, m, _P(mean), _P(var)))
        return mean, var

    def emulate_few(self, pts):
        """emub_predict_few: the latency path for at most 8 points"""
        pts = _c(pts).reshape(-1, self.model.d)
        m = pts.shape[0]
        mean, var = np.empty(m), np.empty(m)
        _check(self.L.emub_predict_few(self.h, _P(pts), self.model.d, m, _P(mean), _P(var)))
        return mean, var

    def emulate_dev(self, d_pts_ptr, m, d_mean_ptr, d_var_ptr):
        _check(self.L.emub_predict_batch_dev(self.h, d_pts_ptr, m, d_mean_ptr, d_var_ptr))

    def beta(self):
        b = np.empty(self.model.p)
        _check(self.L.emub_emulator_beta(self.h, _P(b)))
        return b

    def close(self):
        if getattr(self, "h", None):
            self.L.emub_emulator_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def predict_multi(emulators, pts, training_mean=None, evecs=None, evals=None, few=False):
    """emulate_point_multi (multivar_support.c:103) for a block of points: emulators = the nr PCA-component
    emulators of ONE model.  With the projection data returns (mean, var) of shape (m, nt) in observable space;
    without it the PCA-space values (m, nr) (emulate_point_multi_pca).  few=True: the latency path for <= 8 points."""
    model = emulators[0].model
    fn = model.L.emub_predict_multi_few if few else model.L.emub_predict_multi
    pts = _c(pts).reshape(-1, model.d)
    m, nr = pts.shape[0], len(emulators)
    arr = (_vp * nr)(*[em.h for em in emulators])
    if evecs is not None:
        evecs = _c(evecs).reshape(-1, nr)
        nt = evecs.shape[0]
        mean, var = np.empty((m, nt)), np.empty((m, nt))
        _check(fn(arr, nr, _P(pts), model.d, m, nt, _P(_c(training_mean)), _P(evecs), _P(_c(evals)), _P(mean), _P(var)))
    else:
        mean, var = np.empty((m, nr)), np.empty((m, nr))
        _check(fn(arr, nr, _P(pts), model.d, m, 0, None, None, None, _P(mean), _P(var)))
    return mean, var


# ---- host C layer (madaiemulator_b200/host/libemuhost.so): restart driver over the batched evaluator ----------
HOST_LIB_PATH = os.path.join(_HERE, "host", "libemuhost.so")
HOST_SYMBOLS = ["emub_estimate_default_opts", "emub_sample_scales", "emub_optimization_ranges", "emub_optimization_ranges_ex", "emub_random_init",
                "emub_estimate_thetas", "emub_estimate_thetas_from", "emub_estimate_thetas_multi", "emub_estimate_thetas_multi_devices", "emub_estimate_thetas_multi_devices_ranges", "emub_snapshot_load",
                "emub_snapshot_load_path", "emub_snapshot_free", "emub_multi_emulator_from_snapshot",
                "emub_multi_emulator_destroy", "emub_multi_emulator_predict", "emub_multi_emulator_predict_few", "emub_interactive_stream", "emub_parse_doubles", "emub_fast_strtod", "emub_fast_format17",
                "emub_snapshot_save", "emub_snapshot_save_path", "emub_snapshot_from_arrays"]


class EstimateOpts(ctypes.Structure):
    _fields_ = [("max_tries", _ci), ("nchains", _ci), ("seed", ctypes.c_ulonglong), ("step_size", ctypes.c_double),
                ("tol", ctypes.c_double), ("eps_abs", ctypes.c_double), ("step_max", _ci), ("first_component", _ci),
                ("component_stride", _ci), ("polish_steps", _ci), ("polish_eps", ctypes.c_double), ("value_policy", _ci)]


VALUE_ADAPTIVE, VALUE_ALWAYS_GRADIENT, VALUE_ONLY = 0, 1, 2


class EstimateStats(ctypes.Structure):
    _fields_ = [("evaluations", _ll), ("batches", _ll), ("success_count", _ci), ("finite_count", _ci),
                ("value_evaluations", _ll), ("repeated_points", _ll), ("unused_gradients", _ll)]


def _stats_dict(rc, st):
    return dict(rc=rc, evaluations=st.evaluations, batches=st.batches, success_count=st.success_count,
                finite_count=st.finite_count, value_evaluations=st.value_evaluations, repeated_points=st.repeated_points,
                unused_gradients=st.unused_gradients)


_hostlib = None


def host_lib():
    global _hostlib
    if _hostlib is not None:
        return _hostlib
    if not os.path.exists(HOST_LIB_PATH):
        raise RuntimeError("host layer %s is missing: run `make -C madaiemulator_b200/host`" % HOST_LIB_PATH)
    lib()  # libemub.so first (the host layer links against it)
    H = ctypes.CDLL(HOST_LIB_PATH)
    H.emub_estimate_default_opts.argtypes = [ctypes.POINTER(EstimateOpts)]
    H.emub_estimate_default_opts.restype = None
    H.emub_sample_scales.argtypes = [_dp, _ci, _ci, _ci, _dp]
    H.emub_sample_scales.restype = None
    H.emub_optimization_ranges.argtypes = [_ci, _dp, _ci, _ci, _ci, _dp]
    H.emub_optimization_ranges.restype = None
    H.emub_optimization_ranges_ex.argtypes = [_ci, _dp, _ci, _ci, _ci, _ci, _ci, ctypes.c_double, _dp]
    H.emub_optimization_ranges_ex.restype = None
    H.emub_random_init.argtypes = [ctypes.c_ulonglong, _ci, _dp, _ci, _dp]
    H.emub_random_init.restype = None
    H.emub_estimate_thetas.argtypes = [_vp, _dp, ctypes.POINTER(EstimateOpts), _dp, _dp, ctypes.POINTER(EstimateStats)]
    H.emub_estimate_thetas_from.argtypes = [_vp, _dp, _dp, ctypes.POINTER(EstimateOpts), _dp, _dp, ctypes.POINTER(EstimateStats)]
    H.emub_estimate_thetas_multi.argtypes = [_vp, _ci, _dp, ctypes.POINTER(EstimateOpts), _dp, _dp, ctypes.POINTER(EstimateStats)]
    H.emub_estimate_thetas_multi_devices.argtypes = [_ip, _ci, _dp, _ci, _ci, _ci, _dp, _ci, _ci, _ci, _ci, _ci,
                                                     ctypes.POINTER(EstimateOpts), _dp, _dp, ctypes.POINTER(EstimateStats)]
    _hostlib = H
    return H


def optimization_ranges(kernel, X, use_data_scales=True, fixed_nugget=None):
    """setup_optimization_ranges (optstruct.c:142) on the host: nthetas x 2.  fixed_nugget: the optstruct's
    fixed_nugget_mode = 1 with that value (optstruct.c:217-225)."""
    X = _c(X)
    n, d = X.shape
    nth = d + 2 if kernel == POWEREXP else 3
    r = np.empty((nth, 2))
    host_lib().emub_optimization_ranges_ex(kernel, _P(X), d, n, d, 1 if use_data_scales else 0, 0 if fixed_nugget is None else 1,
                                           0.0 if fixed_nugget is None else float(fixed_nugget), _P(r))
    return r


def random_init(seed, try_index, ranges):
    ranges = _c(ranges)
    x = np.empty(ranges.shape[0])
    host_lib().emub_random_init(seed, try_index, _P(ranges), ranges.shape[0], _P(x))
    return x


def estimate_thetas(model, ranges=None, max_tries=50, nchains=0, seed=1, starts=None, polish_steps=0, value_policy=VALUE_ADAPTIVE,
                    step_max=None):
    """maxWithMultiMin (maxmultimin.c:47) over the batched GPU evaluator.  Returns (thetas[nthetas], best log
    likelihood, stats dict).  starts: optional (max_tries x nthetas) explicit start points.  polish_steps > 0 adds the
    optional refinement run from the best point (emub_estimate.h)."""
    H = host_lib()
    if ranges is None:
        ranges = optimization_ranges(model.kernel, model.X)
    ranges = _c(ranges)
    o = EstimateOpts()
    H.emub_estimate_default_opts(ctypes.byref(o))
    if starts is not None:
        starts = _c(starts).reshape(-1, model.nthetas)
        max_tries = starts.shape[0]
    o.max_tries, o.nchains, o.seed, o.polish_steps, o.value_policy = max_tries, nchains, seed, polish_steps, value_policy
    if step_max is not None:
        o.step_max = step_max
    th = np.zeros(model.nthetas)
    best = ctypes.c_double()
    st = EstimateStats()
    rc = H.emub_estimate_thetas_from(model.h, _P(ranges), _P(starts) if starts is not None else None, ctypes.byref(o),
                                     _P(th), ctypes.byref(best), ctypes.byref(st))
    if rc not in (OK, EDOM):
        _check(rc)
    return th, best.value, _stats_dict(rc, st)


def estimate_thetas_multi(model, ncomp, ranges=None, max_tries=50, nchains=0, seed=1, value_policy=VALUE_ADAPTIVE, step_max=None,
                          first_component=0, component_stride=1):
    """estimate_multi (multivar_support.c:20) with the restart fronts of all ncomp components merged into one batch.
    Returns (thetas[ncomp, nthetas], best log likelihoods[ncomp], stats)."""
    H = host_lib()
    if ranges is None:
        ranges = optimization_ranges(model.kernel, model.X)
    ranges = _c(ranges)
    o = EstimateOpts()
    H.emub_estimate_default_opts(ctypes.byref(o))
    o.max_tries, o.nchains, o.seed, o.value_policy = max_tries, nchains, seed, value_policy
    o.first_component, o.component_stride = first_component, component_stride
    if step_max is not None:
        o.step_max = step_max
    th = np.zeros((ncomp, model.nthetas))
    best = np.zeros(ncomp)
    st = EstimateStats()
    rc = H.emub_estimate_thetas_multi(model.h, ncomp, _P(ranges), ctypes.byref(o), _P(th), _P(best), ctypes.byref(st))
    if rc not in (OK, EDOM):
        _check(rc)
    return th, best, _stats_dict(rc, st)


def estimate_thetas_multi_devices(devices, X, Z, kernel=POWEREXP, order=0, max_tries=50, nchains=0, seed=1, max_slots=0):
    """estimate_multi with the PCA components sharded over several GPUs of one box (component c -> devices[c % ndev]),
    one host thread per device, no exchange between devices.  Returns (thetas[ncomp, nthetas], best[ncomp], stats)."""
    H = host_lib()
    X, Z = _c(X), _c(Z)
    n, d = X.shape
    ncomp = Z.shape[1]
    nth = d + 2 if kernel == POWEREXP else 3
    dev = np.ascontiguousarray(devices, dtype=np.int32)
    o = EstimateOpts()
    H.emub_estimate_default_opts(ctypes.byref(o))
    o.max_tries, o.nchains, o.seed = max_tries, nchains, seed
    th = np.zeros((ncomp, nth))
    best = np.zeros(ncomp)
    st = EstimateStats()
    rc = H.emub_estimate_thetas_multi_devices(dev.ctypes.data_as(_ip), len(dev), _P(X), d, n, d, _P(Z), ncomp, ncomp, kernel, order,
                                              max_slots, ctypes.byref(o), _P(th), _P(best), ctypes.byref(st))
    if rc not in (OK, EDOM):
        _check(rc)
    return th, best, _stats_dict(rc, st)
