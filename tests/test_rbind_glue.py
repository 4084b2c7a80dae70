"""GPU: SURVEY 8f-3 -- the list-shaped R entry points (rbind.c:121-187, :626-724) forwarded to the batched engine,
with the `.C()` calling convention (everything by pointer, matrices column-major)."""
import ctypes
import os

import numpy as np
import pytest

from madaiemulator_b200 import datasets as ds
from tests.helpers import relerr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "madaiemulator_b200", "host", "libemurbind.so")

_dp = ctypes.POINTER(ctypes.c_double)
_ipt = ctypes.POINTER(ctypes.c_int)


def _i(v):
    return ctypes.byref(ctypes.c_int(v))


def _P(a):
    return a.ctypes.data_as(_dp)


def test_list_entry_points_match_per_point_reference():
    from oracle.pyoracle import PortOracle
    L = ctypes.CDLL(LIB)
    n, d, order = 90, 3, 1
    X = ds.synthetic_design(n, d)
    y = ds.synthetic_response(X)
    nth = d + 2
    rng = np.random.default_rng(2)
    B = 9
    # R hands over nthetas columns per row; evalFnMulti uses the first nthetas-1 (nugget, lengths)
    plist = np.column_stack([rng.uniform(-5, -2, B)] + [rng.uniform(0, 1.5, B) for _ in range(d)] + [np.zeros(B)])
    x_cm = np.asfortranarray(X).ravel(order="F").copy()      # column-major n x d, as R stores it
    p_cm = np.asfortranarray(plist).ravel(order="F").copy()  # column-major B x nthetas
    answer = np.zeros(B)
    L.callEvalLhoodList(_P(x_cm), _i(d), _P(p_cm), _i(B), _P(y), _i(n), _i(nth), _P(answer), _i(1), _i(order))
    o = PortOracle(X, y, 1, order)
    for b in range(B):
        assert relerr(answer[b], o.loglik_grad(plist[b, :nth - 1], want_grad=False)["negL"]) < 1e-9
    # emulate at a list of points
    mq = 50
    pts = ds.synthetic_queries(mq, d)
    q_cm = np.asfortranarray(pts).ravel(order="F").copy()
    thetas = np.concatenate([[0.1, -3.5], rng.uniform(0.3, 1.0, d)])
    mean, var = np.zeros(mq), np.zeros(mq)
    L.callEmulateAtList(_P(x_cm), _i(d), _P(q_cm), _i(mq), _P(y), _i(n), _P(thetas), _i(nth), _P(mean), _P(var), _i(1), _i(order))
    m_ref, v_ref = o.emulator(thetas).emulate(pts)
    assert relerr(mean, m_ref, 1e-3) < 1e-9
    assert np.max(np.abs(var - v_ref)) < 1e-9 * 2.0
    L.rbind_glue_reset()


REF_RBIND = os.path.join(ROOT, "oracle", "_ref", "librbind_ref.so")        # the reference's rbind.c on its own CPU path
DROPIN_RBIND = os.path.join(ROOT, "oracle", "_ref", "librbind_dropin.so")  # the same file, unmodified, on the engine


def _mc_case(kernel, order, n=90, d=3, ny=3):
    X = ds.synthetic_design(n, d, seed=ds.SEED + 100 + kernel)
    Y = np.stack([ds.synthetic_response(X, t) for t in range(ny)], axis=1)
    rng = np.random.default_rng(10 * kernel + order)
    if kernel == 1:
        TH = np.stack([np.concatenate([[rng.uniform(-0.5, 0.5), rng.uniform(-4.5, -3)], rng.uniform(0.3, 1.0, d)]) for _ in range(ny)])
    else:  # Matern: amplitude and nugget raw (emulator.c:355-356)
        TH = np.stack([np.array([rng.uniform(0.8, 1.5), rng.uniform(0.01, 0.05), rng.uniform(0.2, 0.8)]) for _ in range(ny)])
    pts = ds.synthetic_queries(25, d)
    pts[3] = X[11]  # on a design point
    return X, Y, TH, pts


def _run_mc(L, X, Y, TH, pts, kernel, order):
    """setupEmulateMC / callEmulateMC and the Multi twins through one library; returns (uni mean/var, multi mean/var)"""
    n, d = X.shape
    ny, nth = TH.shape
    x_cm = np.asfortranarray(X).ravel(order="F").copy()
    y0 = np.ascontiguousarray(Y[:, 0])
    th0 = np.ascontiguousarray(TH[0])
    L.setupEmulateMC(_P(x_cm), _i(d), _P(y0), _i(n), _P(th0), _i(nth), _i(kernel), _i(order))
    um, uv = np.zeros(len(pts)), np.zeros(len(pts))
    for q, p in enumerate(pts):
        m_, v_ = ctypes.c_double(), ctypes.c_double()
        pp = np.ascontiguousarray(p)
        L.callEmulateMC(_P(pp), ctypes.byref(m_), ctypes.byref(v_))
        um[q], uv[q] = m_.value, v_.value
    L.freeEmulateMC()
    y_cm = np.asfortranarray(Y).ravel(order="F").copy()
    th_cm = np.asfortranarray(TH).ravel(order="F").copy()
    L.setupEmulateMCMulti(_P(x_cm), _i(d), _P(y_cm), _i(ny), _i(n), _P(th_cm), _i(nth), _i(kernel), _i(order))
    mm, mv = np.zeros((len(pts), ny)), np.zeros((len(pts), ny))
    for q, p in enumerate(pts):
        pp = np.ascontiguousarray(p)
        a, b = np.zeros(ny), np.zeros(ny)
        L.callEmulateMCMulti(_P(pp), _i(ny), _P(a), _P(b))
        mm[q], mv[q] = a, b
    L.freeEmulateMCMulti(_i(ny))
    return um, uv, mm, mv


def _kappa(TH, kernel):
    return float(np.max(np.exp(TH[:, 0]) + np.exp(TH[:, 1]))) if kernel == 1 else float(np.max(TH[:, 0] + TH[:, 1]))


@pytest.mark.parametrize("kernel,order", [(1, 0), (1, 2), (2, 1), (3, 0)])
def test_monte_carlo_entry_points_match_the_reference_rbind(kernel, order):
    """setupEmulateMC / callEmulateMC / setupEmulateMCMulti / callEmulateMCMulti (rbind.c:299-590): the handle-based
    engine versions against the reference's own rbind.c compiled unmodified (host-side C^-1 + emulateQuick per point)."""
    if not os.path.exists(REF_RBIND):
        pytest.skip("oracle/_ref/librbind_ref.so not built")
    X, Y, TH, pts = _mc_case(kernel, order)
    ref = _run_mc(ctypes.CDLL(REF_RBIND), X, Y, TH, pts, kernel, order)
    L = ctypes.CDLL(LIB)
    got = _run_mc(L, X, Y, TH, pts, kernel, order)
    kap = _kappa(TH, kernel)
    for g, r, is_var in zip(got, ref, (False, True, False, True)):
        if is_var:
            assert np.max(np.abs(g - r)) < 1e-9 * max(1.0, kap)
        else:
            assert relerr(g, r, 1e-3) < 1e-9
    # a second set-up replaces the first one; a call after free is refused rather than answered from stale data
    got2 = _run_mc(L, X, Y[:, ::-1].copy(), TH[::-1].copy(), pts, kernel, order)
    assert np.max(np.abs(got2[2][:, ::-1] - got[2])) < 1e-12 * max(1.0, np.max(np.abs(got[2])))
    L.rbind_glue_reset()


@pytest.mark.parametrize("kernel,order", [(1, 1), (3, 0)])
def test_unmodified_rbind_runs_on_the_engine(kernel, order):
    """The reference's rbind.c, unchanged, linked with integration/libemu_glue.c: its set-up calls makeCovMatrix and
    chol_inverse_cov_matrix (now the engine's), every callEmulateMC[Multi] goes through emulateQuick (now the cached
    factor + the latency path), callEvalLhoodList / callEmulateAtList / callEmulateAtPt through evalFnMulti and
    emulateAtPoint[List]."""
    if not (os.path.exists(REF_RBIND) and os.path.exists(DROPIN_RBIND)):
        pytest.skip("oracle/_ref/librbind_*.so not built")
    X, Y, TH, pts = _mc_case(kernel, order)
    R = ctypes.CDLL(REF_RBIND)
    D = ctypes.CDLL(DROPIN_RBIND)
    ref = _run_mc(R, X, Y, TH, pts, kernel, order)
    got = _run_mc(D, X, Y, TH, pts, kernel, order)
    kap = _kappa(TH, kernel)
    for g, r, is_var in zip(got, ref, (False, True, False, True)):
        if is_var:
            assert np.max(np.abs(g - r)) < 1e-9 * max(1.0, kap)
        else:
            assert relerr(g, r, 1e-3) < 1e-9
    n, d = X.shape
    nth = TH.shape[1]
    x_cm = np.asfortranarray(X).ravel(order="F").copy()
    y0 = np.ascontiguousarray(Y[:, 0])
    th0 = np.ascontiguousarray(TH[0])
    # callEmulateAtPt (rbind.c:207-270) and callEmulateAtList
    for lib_ in (R, D):
        lib_.out = []
        for p in pts[:4]:
            m_, v_ = ctypes.c_double(), ctypes.c_double()
            pp = np.ascontiguousarray(p)
            lib_.callEmulateAtPt(_P(x_cm), _i(d), _P(pp), _P(y0), _i(n), _P(th0), _i(nth), ctypes.byref(m_), ctypes.byref(v_), _i(kernel), _i(order))
            lib_.out.append((m_.value, v_.value))
    a, b = np.array(R.out), np.array(D.out)
    assert relerr(b[:, 0], a[:, 0], 1e-3) < 1e-9 and np.max(np.abs(b[:, 1] - a[:, 1])) < 1e-9 * max(1.0, kap)
    q_cm = np.asfortranarray(pts).ravel(order="F").copy()
    outs = []
    for lib_ in (R, D):
        mean, var = np.zeros(len(pts)), np.zeros(len(pts))
        lib_.callEmulateAtList(_P(x_cm), _i(d), _P(q_cm), _i(len(pts)), _P(y0), _i(n), _P(th0), _i(nth), _P(mean), _P(var), _i(kernel), _i(order))
        outs.append((mean, var))
    assert relerr(outs[1][0], outs[0][0], 1e-3) < 1e-9 and np.max(np.abs(outs[1][1] - outs[0][1])) < 1e-9 * max(1.0, kap)
    if kernel == 1:
        # likelihood over a list of thetas (rbind.c:626-724); Matern training is non-functional in the reference (Q6)
        rng = np.random.default_rng(3)
        B = 5
        plist = np.column_stack([rng.uniform(-5, -2, B)] + [rng.uniform(0, 1.5, B) for _ in range(d)] + [np.zeros(B)])
        p_cm = np.asfortranarray(plist).ravel(order="F").copy()
        ans = []
        for lib_ in (R, D):
            answer = np.zeros(B)
            lib_.callEvalLhoodList(_P(x_cm), _i(d), _P(p_cm), _i(B), _P(y0), _i(n), _i(nth), _P(answer), _i(kernel), _i(order))
            ans.append(answer)
        assert relerr(ans[1], ans[0]) < 1e-9
    D.libemu_glue_reset()
