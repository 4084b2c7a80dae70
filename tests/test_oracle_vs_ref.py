"""CPU: pin the restatement oracle against the reference's own sources compiled here
(oracle/_ref/libemu_ref.so).  Skipped where that library is absent."""
import numpy as np
import pytest

from madaiemulator_b200 import datasets as ds
from oracle.pyoracle import DET_LOGSUM, DET_PRODUCT, PortOracle, RefOracle, ref_available
from tests.helpers import relerr

pytestmark = pytest.mark.skipif(not ref_available(), reason="oracle/_ref/libemu_ref.so not built")


@pytest.mark.parametrize("n,d,order", [(40, 1, 0), (64, 3, 1), (96, 6, 2), (130, 10, 3), (300, 15, 1)])
def test_powerexp_value_grad_prediction(n, d, order):
    X = ds.synthetic_design(n, d, seed=ds.SEED + n)
    y = ds.synthetic_response(X, seed=ds.SEED + n)
    r, p = RefOracle(X, y, 1, order), PortOracle(X, y, 1, order)
    rng = np.random.default_rng(n)
    for _ in range(2):
        th = np.concatenate([[rng.uniform(-5, -2)], rng.uniform(0.0, 1.5, d)])
        full = np.concatenate([[rng.uniform(-1, 1)], th])
        assert np.array_equal(r.cov_matrix(full), p.cov_matrix(full))
        a, b = r.eval_logsum(th), p.loglik_grad(th, DET_LOGSUM)
        assert a["status"] == b["status"] == 0
        assert relerr(b["negL"], a["negL"]) < 1e-12
        assert relerr(b["sigma2"], a["sigma2"]) < 1e-10
        assert relerr(b["grad"], r.grad(th), 1e-8) < 1e-9
        lit = p.loglik_grad(th, DET_PRODUCT, want_grad=False)["negL"]
        assert relerr(lit, r.eval(th)) < 1e-12
        pts = ds.synthetic_queries(20, d, seed=n)
        pts[0] = X[n // 2]
        m1, v1 = r.emulator(full).emulate(pts)
        m2, v2 = p.emulator(full).emulate(pts)
        assert np.max(np.abs(m1 - m2)) < 1e-11
        assert np.max(np.abs(v1 - v2)) < 1e-11


@pytest.mark.parametrize("kernel", [2, 3])
def test_matern_function_level(kernel):
    X = ds.synthetic_design(50, 4)
    y = ds.synthetic_response(X)
    r, p = RefOracle(X, y, kernel, 1), PortOracle(X, y, kernel, 1)
    full = np.array([1.7, 0.05, 0.4])
    assert np.array_equal(r.cov_matrix(full), p.cov_matrix(full))
    pts = ds.synthetic_queries(10, 4)
    pts[3] = X[9]
    m1, v1 = r.emulator(full).emulate(pts)
    m2, v2 = p.emulator(full).emulate(pts)
    assert np.max(np.abs(m1 - m2)) < 1e-11
    assert np.max(np.abs(v1 - v2)) < 1e-11


def test_reference_matern_training_is_nonfunctional():
    """Q6: evalFnMulti passes amp = 0 (raw) to the Matern kernels -> C = theta_1 * delta with
    theta_1 in [-5,-2] -> never positive definite -> NaN.  Documents why deviation D-2 exists."""
    X = ds.synthetic_design(30, 2)
    y = ds.synthetic_response(X)
    r = RefOracle(X, y, 2, 0)
    assert np.isnan(r.eval(np.array([-3.0, 0.2])))


def test_random_inits_follow_ranges():
    X = ds.synthetic_design(64, 3)
    y = ds.synthetic_response(X)
    r = RefOracle(X, y, 1, 0)
    rg = r.ranges()
    x0 = r.random_inits(7, 50)
    assert np.all(x0 >= rg[:, 0]) and np.all(x0 <= rg[:, 1])


def test_host_ranges_with_fixed_nugget_and_default_scales():
    """emub_optimization_ranges_ex against the reference's setup_optimization_ranges (optstruct.c:142-226) with
    fixed_nugget_mode = 1 (:217-225) and use_data_scales = 0 (:205-211), both kernels."""
    from madaiemulator_b200 import datasets as ds
    from madaiemulator_b200 import engine
    from oracle.pyoracle import RefOracle, ref_available
    if not ref_available():
        pytest.skip("oracle/_ref not built")
    X = ds.synthetic_design(60, 4)
    y = ds.synthetic_response(X)
    for kernel in (1, 2, 3):
        ref = RefOracle(X, y, kernel, 0)
        for uds in (True, False):
            for fn in (None, 0.001, -3.5):
                got = engine.optimization_ranges(kernel, X, use_data_scales=uds, fixed_nugget=fn)
                assert np.array_equal(got, ref.ranges_ex(uds, fn)), (kernel, uds, fn)
        assert np.array_equal(engine.optimization_ranges(kernel, X), ref.ranges())
