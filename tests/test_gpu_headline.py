"""GPU parity at the BASELINE sizes against references the GPU did not compute.

* headline / cfg5 shape (n=4096, d=10, power-exponential, order 0): -L, sigma^2, all 11 gradient components and
  (mean, variance) on 1000 query points against the CPU port oracle (oracle/emu_oracle.c, pinned to the reference
  build), i.e. evalFnMulti / gradFnMulti (maxmultimin.c:288-394, :416-608) and emulate_point (emulator_struct.c:124-143).
* cfg4 shape (n=8192, d=15): the oracle's covariance matrix, then host LAPACK (numpy) for the inverse, the
  regression algebra and the WHOLE gradient vector of the literal formula (Appendix A of SURVEY.md).

Tolerances are the rules of tests/test_gpu_parity.py: 1e-9 relative; cancelling quantities (sigma^2, gradient
components, variance) relative to the magnitude of their constituents.
"""
import threading

import numpy as np
import pytest

from madaiemulator_b200 import datasets as ds
from tests.helpers import relerr

pytestmark = pytest.mark.gpu

TOL = 1e-9


@pytest.fixture(scope="module")
def ctx():
    from madaiemulator_b200 import engine
    c = engine.Context(0)
    yield c
    c.close()


def _grad_err(g, gref):
    scale = np.maximum(np.abs(gref), 1e-3 * np.max(np.abs(gref)) + 1e-300)
    return float(np.max(np.abs(g - gref) / scale))


def test_headline_n4096_d10_against_the_port_oracle(ctx):
    from madaiemulator_b200 import engine
    from oracle.pyoracle import PortOracle
    n, d, order = 4096, 10, 0
    X = ds.synthetic_design(n, d)
    y = ds.synthetic_response(X)
    # a point of the optimiser's search box with unequal lengths (the bench draws its batch from the same box)
    rng = np.random.default_rng(4096)
    th = np.concatenate([[-3.7], rng.uniform(0.4, 1.6, d)])
    o = PortOracle(X, y, 1, order)
    full_holder, ref_holder, emu_holder = {}, {}, {}
    pts = ds.synthetic_queries(1000, d)
    pts[0] = X[n // 2]   # a query on a design point: coincidence nugget in k and kappa (emulator.c:136-150)
    pts[1] = X[n - 1]

    # the two CPU legs (about a minute each) run side by side; ctypes releases the GIL
    def leg_lik():
        ref_holder["r"] = o.loglik_grad(th)

    t1 = threading.Thread(target=leg_lik)
    t1.start()
    m = engine.Model(ctx, X, y, 1, order, max_slots=2)
    r = m.loglik_grad(th)
    assert r["status"] == 0
    full = np.concatenate([[np.log(r["sigma2"])], th])  # what estimate_thetas hands to alloc_emulator_struct (maxmultimin.c:757-769)

    def leg_emu():
        e = o.emulator(full)
        emu_holder["mv"] = e.emulate(pts)
        emu_holder["beta"] = e.beta()

    t2 = threading.Thread(target=leg_emu)
    t2.start()
    e = m.emulator(full)
    mean, var = e.emulate(pts)
    mean_few, var_few = e.emulate_few(pts[:8])
    beta = e.beta()
    r0 = m.loglik_grad_batch(th[None, :], want_grad=False)
    t1.join()
    t2.join()
    ref = ref_holder["r"]
    assert ref["status"] == 0
    assert relerr(r["negL"], ref["negL"]) < TOL
    assert relerr(r["logdet"], ref["logdet"]) < TOL
    assert relerr(r["beta"], ref["beta"], 1e-6) < TOL
    C = o.cov_matrix(np.concatenate([[0.0], th]))
    s2scale = abs(float(y @ np.linalg.solve(C, y))) / n
    assert abs(r["sigma2"] - ref["sigma2"]) < TOL * s2scale
    assert r["grad"].shape == (d + 1,)
    assert _grad_err(r["grad"], ref["grad"]) < TOL, (r["grad"], ref["grad"])
    # the value-only call returns the same bits
    assert r0["negL"][0] == r["negL"] and r0["sigma2"][0] == r["sigma2"]
    m_ref, v_ref = emu_holder["mv"]
    kappa = np.exp(full[0]) + np.exp(full[1])
    assert relerr(mean, m_ref, 1e-3) < TOL
    assert np.max(np.abs(var - v_ref)) < TOL * max(1.0, kappa)
    assert relerr(beta, emu_holder["beta"], 1e-6) < TOL
    assert relerr(mean_few, m_ref[:8], 1e-3) < TOL
    assert np.max(np.abs(var_few - v_ref[:8])) < TOL * max(1.0, kappa)
    # the design-point queries reproduce the training data up to the nugget's share
    assert abs(mean[0] - y[n // 2]) < 0.1 and var[0] < 0.2 * kappa
    e.close()
    m.close()


def test_cfg4_n8192_d15_against_host_lapack(ctx):
    """BASELINE config 4 shape.  Reference values: the oracle's C (emuo_cov_matrix, emulator.c:636) and float64 LAPACK
    on the host for everything after it; nothing on the reference side of an assert comes from the GPU."""
    from madaiemulator_b200 import engine
    from oracle.pyoracle import PortOracle
    n, d = 8192, 15
    X = ds.synthetic_design(n, d)
    y = ds.synthetic_response(X)
    rng = np.random.default_rng(8192)
    th = np.concatenate([[-4.0], rng.uniform(0.7, 1.5, d)])
    m = engine.Model(ctx, X, y, 1, 0, max_slots=2)
    r = m.loglik_grad_batch(np.stack([th, th + 0.01]))
    assert np.all(r["status"] == 0)
    r1 = m.loglik_grad_batch(th[None, :])
    assert r1["negL"][0] == r["negL"][0] and np.array_equal(r1["grad"][0], r["grad"][0])

    o = PortOracle(X, y, 1, 0)
    C = o.cov_matrix(np.concatenate([[0.0], th]))
    assert relerr(m.cov_matrix(np.concatenate([[0.0], th]))[::97], C[::97], 1e-300) < TOL
    Lc = np.linalg.cholesky(C)
    logdet = 2.0 * np.sum(np.log(np.diag(Lc)))
    Cinv = np.linalg.inv(C)
    Cinv = 0.5 * (Cinv + Cinv.T)
    del Lc
    H = np.ones(n)
    a = Cinv @ y                    # alpha = C^-1 y (raw y, maxmultimin.c:594)
    b = Cinv @ H
    beta = (H @ a) / (H @ b)        # regression.c:120-176, p = 1
    res = y - H * beta
    negL = 0.5 * logdet + (n / 2.0) * 1.83788 + 0.5 * (res @ (Cinv @ res))   # estimator-fns.c:48,85-95
    sigma2 = (y @ (Cinv @ res)) / n                                           # maxmultimin.c:259-263
    assert relerr(r["negL"][0], negL) < TOL
    assert abs(r["sigma2"][0] - sigma2) < TOL * abs(y @ a) / n
    # gradient, literal formula (maxmultimin.c:514-538, 583-602; emulator.c:181,203), every component
    nug = np.exp(th[0])
    g = np.zeros(d + 1)
    terms = np.zeros(d + 1)
    t_tr, t_q = -0.5 * nug * np.trace(Cinv), 0.5 * nug * (a @ a)
    g[0] = -1.0 * (t_tr + t_q)
    terms[0] = abs(t_tr) + abs(t_q)
    for k in range(d):
        dl = X[:, k][:, None] - X[:, k][None, :]
        q = dl * dl
        D = np.exp(-0.5 * np.exp(-2.0 * th[1 + k]) * q - 2.0 * th[1 + k]) * q
        t_tr = -0.5 * sigma2 * np.sum(Cinv * D)
        t_q = 0.5 * sigma2 * (a @ (D @ a))
        g[1 + k] = -1.0 * (t_tr + t_q)
        terms[1 + k] = abs(t_tr) + abs(t_q)
        del dl, q, D
    err = np.abs(r["grad"][0] - g) / terms
    assert np.max(err) < TOL, (r["grad"][0], g, err)
    # prediction at the estimated amplitude: 256 points incl. a design point against C^-1 algebra on the host
    full = np.concatenate([[np.log(sigma2)], th])
    amp, kappa = np.exp(full[0]), np.exp(full[0]) + nug
    Cf = amp * (C - nug * np.eye(n)) + nug * np.eye(n)   # C(theta_full) = amp c + nug delta from the unit-amplitude matrix
    pts = ds.synthetic_queries(256, d)
    pts[0] = X[17]
    K = np.stack([np.array([o.cov_pair(X[i], pts[qi], full) for i in range(n)]) for qi in range(8)], axis=1)
    K[K < 1e-10] = 0.0
    Cfi_K = np.linalg.solve(Cf, K)
    Cfi_y = np.linalg.solve(Cf, np.stack([y, H], axis=1))
    beta_f = (H @ Cfi_y[:, 0]) / (H @ Cfi_y[:, 1])
    mean_ref = beta_f + K.T @ (Cfi_y[:, 0] - beta_f * Cfi_y[:, 1])
    rho = 1.0 - K.T @ Cfi_y[:, 1]
    var_ref = kappa - np.sum(K * Cfi_K, axis=0) + rho * rho / (H @ Cfi_y[:, 1])
    e = m.emulator(full)
    mean, var = e.emulate(pts)
    assert relerr(mean[:8], mean_ref, 1e-3) < TOL
    assert np.max(np.abs(var[:8] - var_ref)) < TOL * max(1.0, kappa)
    e.close()
    m.close()
