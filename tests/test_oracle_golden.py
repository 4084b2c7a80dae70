"""CPU: the plain-C restatement oracle (oracle/emu_oracle.c) against the golden fixtures that the
reference's own sources produced (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from oracle.pyoracle import DET_LOGSUM, DET_PRODUCT, PortOracle
from tests.helpers import golden_names, load_golden, relerr

NAMES = golden_names()


@pytest.mark.parametrize("name", NAMES)
def test_port_oracle_matches_golden(name):
    c = load_golden(name)
    o = PortOracle(c["X"], c["y"], c["kernel"], c["order"])
    C = o.cov_matrix(c["theta_full"])
    # covariance entries: identical arithmetic order -> bit exact
    assert np.array_equal(C[c["cov_rows"]], c["cov_row_values"].reshape(len(c["cov_rows"]), -1))
    assert np.array_equal(np.diag(C), c["cov_diag"])
    assert np.array_equal(o.h_matrix().ravel(), c["H"].ravel())
    assert np.array_equal(o.ranges().ravel(), c["ranges"].ravel())
    e = o.emulator(c["theta_full"])
    mean, var = e.emulate(c["pts"])
    assert relerr(mean, c["emu_mean"], 1e-3) < 1e-12
    assert np.max(np.abs(var - c["emu_var"])) < 1e-12 * max(1.0, float(c["kappa"]))
    assert relerr(e.beta(), c["emu_beta"], 1e-6) < 1e-10
    if c["kernel"] == 1:
        r = o.loglik_grad(c["theta_less_amp"], DET_LOGSUM)
        assert r["status"] == 0
        assert relerr(r["negL"], c["negL_logsum"]) < 1e-13
        assert relerr(r["logdet"], c["logdet"]) < 1e-13
        assert relerr(r["sigma2"], c["sigma2"]) < 1e-11
        assert relerr(r["beta"], c["beta"], 1e-6) < 1e-10
        assert relerr(r["grad"], c["grad"], 1e-6) < 1e-10
        lit = o.loglik_grad(c["theta_less_amp"], DET_PRODUCT, want_grad=False)
        assert relerr(lit["negL"], c["negL_literal"]) < 1e-13
        D = o.deriv_matrix(c["theta_less_amp"][1], 2)
        assert np.array_equal(D[0], c["deriv2_row0"])


def test_logsum_equals_product_where_reference_is_finite():
    """Deviation D-1: sum 2 log L_ii == log((prod L_ii)^2) wherever the reference's product is finite."""
    for name in ("uni-simple-o1", "multi-simple-pc0-o0", "synthetic-n256-d10-o1"):
        c = load_golden(name)
        assert np.isfinite(c["negL_literal"])
        assert relerr(c["negL_logsum"], c["negL_literal"]) < 1e-12


def test_product_determinant_underflows_at_scale():
    """Q3: at n ~ 1000 the literal product underflows -> the reference's -L is -inf/+inf."""
    from madaiemulator_b200 import datasets as ds
    X = ds.synthetic_design(1100, 10)
    y = ds.synthetic_response(X)
    o = PortOracle(X, y, 1, 0)
    th = ds.default_theta_less_amp(10)
    lit = o.loglik_grad(th, DET_PRODUCT, want_grad=False)
    ok = o.loglik_grad(th, DET_LOGSUM, want_grad=False)
    assert not np.isfinite(lit["negL"])
    assert np.isfinite(ok["negL"])


def test_nonpd_reports_nan():
    X = np.array([[0.0], [0.0], [1.0]])  # duplicated design point, tiny nugget -> still PD thanks to nugget
    y = np.array([1.0, 1.1, 0.3])
    o = PortOracle(X, y, 1, 0)
    r = o.loglik_grad(np.array([-40.0, 0.0]), want_grad=False)
    # exp(-40) nugget on a duplicated row: numerically singular -> status 1 and NaN, as evalFnMulti returns
    assert r["status"] in (0, 1)
    if r["status"] == 1:
        assert np.isnan(r["negL"])
