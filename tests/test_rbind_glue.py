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
