"""CPU: the optimiser of the restart driver (madaiemulator_b200/host/emub_bfgs.c, a restatement of the published
vector-BFGS + Fletcher line search the reference selects at maxmultimin.c:683 -- GSL is an un-vendored dependency, so
there is no third-party build to compare with; the checks are the algorithm's own guarantees on classic functions)."""
import ctypes
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_dp = ctypes.POINTER(ctypes.c_double)
F_T = ctypes.CFUNCTYPE(ctypes.c_double, _dp, ctypes.c_void_p)
DF_T = ctypes.CFUNCTYPE(None, _dp, ctypes.c_void_p, _dp)
FDF_T = ctypes.CFUNCTYPE(None, _dp, ctypes.c_void_p, _dp, _dp)


class Fn(ctypes.Structure):
    _fields_ = [("n", ctypes.c_size_t), ("f", F_T), ("df", DF_T), ("fdf", FDF_T), ("ctx", ctypes.c_void_p)]


@pytest.fixture(scope="module")
def lib():
    L = ctypes.CDLL(os.path.join(ROOT, "madaiemulator_b200", "host", "libemuhost.so"))
    L.emub_bfgs_alloc.restype = ctypes.c_void_p
    L.emub_bfgs_alloc.argtypes = [ctypes.c_size_t]
    L.emub_bfgs_free.argtypes = [ctypes.c_void_p]
    L.emub_bfgs_set.argtypes = [ctypes.c_void_p, ctypes.POINTER(Fn), _dp, ctypes.c_double, ctypes.c_double]
    L.emub_bfgs_iterate.argtypes = [ctypes.c_void_p]
    for f in ("emub_bfgs_x", "emub_bfgs_gradient", "emub_bfgs_dx"):
        getattr(L, f).restype = _dp
        getattr(L, f).argtypes = [ctypes.c_void_p]
    L.emub_bfgs_minimum.restype = ctypes.c_double
    L.emub_bfgs_minimum.argtypes = [ctypes.c_void_p]
    L.emub_bfgs_test_gradient.argtypes = [_dp, ctypes.c_size_t, ctypes.c_double]
    return L


class Problem:
    def __init__(self, n, f, g):
        self.n, self.f, self.g = n, f, g
        self.nf = self.ng = 0
        self.points = []

        def cf(x, _):
            v = np.ctypeslib.as_array(x, (n,)).copy()
            self.nf += 1
            self.points.append(("f", v))
            return float(f(v))

        def cdf(x, _, out):
            v = np.ctypeslib.as_array(x, (n,)).copy()
            self.ng += 1
            self.points.append(("g", v))
            np.ctypeslib.as_array(out, (n,))[:] = g(v)

        def cfdf(x, _, fo, out):
            v = np.ctypeslib.as_array(x, (n,)).copy()
            self.nf += 1
            self.ng += 1
            self.points.append(("fg", v))
            fo[0] = float(f(v))
            np.ctypeslib.as_array(out, (n,))[:] = g(v)

        self._keep = (F_T(cf), DF_T(cdf), FDF_T(cfdf))
        self.fn = Fn(n, self._keep[0], self._keep[1], self._keep[2], None)


def rosen(x):
    return 100.0 * (x[1] - x[0] ** 2) ** 2 + (1.0 - x[0]) ** 2


def rosen_g(x):
    return np.array([-400.0 * x[0] * (x[1] - x[0] ** 2) - 2.0 * (1.0 - x[0]), 200.0 * (x[1] - x[0] ** 2)])


def _run(L, pb, x0, step, tol, eps, max_iter):
    s = L.emub_bfgs_alloc(pb.n)
    x0 = np.array(x0, dtype=np.float64)
    assert L.emub_bfgs_set(s, ctypes.byref(pb.fn), x0.ctypes.data_as(_dp), step, tol) == 0
    hist = [(x0.copy(), L.emub_bfgs_minimum(s), np.ctypeslib.as_array(L.emub_bfgs_gradient(s), (pb.n,)).copy())]
    status = None
    for _ in range(max_iter):
        status = L.emub_bfgs_iterate(s)
        if status == 27:  # EMUB_BFGS_ENOPROG
            break
        x = np.ctypeslib.as_array(L.emub_bfgs_x(s), (pb.n,)).copy()
        g = np.ctypeslib.as_array(L.emub_bfgs_gradient(s), (pb.n,)).copy()
        hist.append((x, L.emub_bfgs_minimum(s), g))
        status = L.emub_bfgs_test_gradient(g.ctypes.data_as(_dp), pb.n, eps)
        if status == 0:
            break
    L.emub_bfgs_free(s)
    return hist, status


def test_rosenbrock_converges_and_every_step_is_a_wolfe_step(lib):
    pb = Problem(2, rosen, rosen_g)
    tol = 0.1
    hist, status = _run(lib, pb, [-1.2, 1.0], 0.1, tol, 1e-6, 200)
    assert status == 0 and np.allclose(hist[-1][0], [1.0, 1.0], atol=1e-6)
    assert len(hist) < 80
    for (x0, f0, g0), (x1, f1, g1) in zip(hist, hist[1:]):
        assert f1 == rosen(x1) and np.array_equal(g1, rosen_g(x1))  # the state is what the callbacks returned at x
        p = x1 - x0
        slope0, slope1 = g0 @ p, g1 @ p
        assert slope0 < 0                                 # a descent direction
        assert f1 <= f0 + 0.01 * slope0 + 1e-14 * abs(f0)  # sufficient decrease, rho = 0.01
        assert abs(slope1) <= tol * abs(slope0) * (1 + 1e-9) + 1e-14  # curvature condition with sigma = tol


def test_quadratic_terminates_like_conjugate_directions(lib):
    """on a strictly convex quadratic, BFGS with an accurate line search finds the minimum in at most n iterations"""
    rng = np.random.default_rng(3)
    n = 6
    A = rng.standard_normal((n, n))
    A = A @ A.T + n * np.eye(n)
    b = rng.standard_normal(n)
    pb = Problem(n, lambda x: 0.5 * x @ A @ x - b @ x, lambda x: A @ x - b)
    hist, status = _run(lib, pb, np.zeros(n), 1.0, 1e-10, 1e-7, 50)
    assert status == 0 and len(hist) - 1 <= n + 1
    assert np.allclose(hist[-1][0], np.linalg.solve(A, b), atol=1e-7)


def test_reference_settings_stop_rule_and_no_progress(lib):
    """step 1.5, tol 0.5, |g| < 0.1, <= 30 iterations (maxmultimin.c:644-656): a smooth bowl stops by the gradient test; started AT
    a minimum the first iteration reports no progress (ENOPROG, what maxmultimin.c:704-708 checks for)"""
    w = np.array([1.0, 1.5, 2.0, 2.5])
    a = np.array([-3.0, -2.8, -2.6, -2.4])
    f = lambda x: float(np.sum(w * (x - a) ** 2 + 0.1 * (x - a) ** 4))
    g = lambda x: 2.0 * w * (x - a) + 0.4 * (x - a) ** 3
    pb = Problem(4, f, g)
    hist, status = _run(lib, pb, [0.5, -4.0, 1.0, -1.0], 1.5, 0.5, 0.1, 30)
    assert status == 0 and np.linalg.norm(hist[-1][2]) < 0.1
    fs = [h[1] for h in hist]
    assert all(f1 <= f0 for f0, f1 in zip(fs, fs[1:]))
    # the optimiser asks for values (line-search trial points) and gradients separately: that split is what the
    # value-only path of the engine serves
    kinds = [k for k, _ in pb.points]
    assert kinds[0] == "fg" and "f" in kinds and "g" in kinds
    pb2 = Problem(4, f, g)
    hist2, status2 = _run(lib, pb2, a, 1.5, 0.5, 0.1, 30)
    assert status2 == 27 and len(hist2) == 1


def test_nan_objective_does_not_hang(lib):
    """a not-positive-definite covariance matrix comes back as NaN (maxmultimin.c:327-350): the line search must give
    up, not loop"""
    def f(x):
        return float("nan") if x[0] > 0.5 else float((x[0] - 1.0) ** 2 + x[1] ** 2)

    def g(x):
        return np.array([np.nan, np.nan]) if x[0] > 0.5 else np.array([2.0 * (x[0] - 1.0), 2.0 * x[1]])

    pb = Problem(2, f, g)
    hist, status = _run(lib, pb, [0.0, 0.3], 1.5, 0.5, 0.1, 30)
    assert pb.nf + pb.ng < 2000
    assert np.all(np.isfinite(hist[0][0]))
