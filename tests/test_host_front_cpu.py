"""CPU: the host restart driver (madaiemulator_b200/host/emub_estimate.c + emub_bfgs.c: evaluation fronts, value
policy, restart bookkeeping, component sharding over devices) against a mock of the C-ABI entry points it calls
(tests/mock/mock_emub.c: an analytic objective with a known minimum per component).  No GPU, no CUDA library."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "madaiemulator_b200", "host")

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("mock") / "libhostmock.so")
    subprocess.check_call(["gcc", "-std=gnu99", "-O2", "-fPIC", "-shared", "-o", out, os.path.join(HOST, "emub_estimate.c"),
                           os.path.join(HOST, "emub_bfgs.c"), os.path.join(ROOT, "tests", "mock", "mock_emub.c"), "-lm", "-lpthread"])
    L = ctypes.CDLL(out)
    from madaiemulator_b200.engine import EstimateOpts, EstimateStats
    L.emub_estimate_default_opts.argtypes = [ctypes.POINTER(EstimateOpts)]
    L.emub_estimate_thetas_multi.argtypes = [ctypes.c_void_p, ctypes.c_int, _dp, ctypes.POINTER(EstimateOpts), _dp, _dp, ctypes.POINTER(EstimateStats)]
    L.emub_estimate_thetas_multi_devices_ranges.argtypes = [_ip, ctypes.c_int, _dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _dp, ctypes.c_int,
                                                            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _dp,
                                                            ctypes.POINTER(EstimateOpts), _dp, _dp, ctypes.POINTER(EstimateStats)]
    L.emub_ctx_create.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
    L.emub_model_create.argtypes = [ctypes.c_void_p, _dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _dp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                    ctypes.POINTER(ctypes.c_void_p)]
    L.emub_model_set_training_multi.argtypes = [ctypes.c_void_p, _dp, ctypes.c_int, ctypes.c_int]
    L.emub_model_destroy.argtypes = [ctypes.c_void_p]
    for f in ("mock_calls", "mock_points", "mock_value_points", "mock_max_batch"):
        getattr(L, f).restype = ctypes.c_longlong
    return L


def _P(a):
    return a.ctypes.data_as(_dp)


def _target(c, d):
    return np.array([-3.0 + 0.37 * c + 0.21 * i for i in range(d + 1)])


def _run(L, d, ncomp, policy, tries=6, chains=6, seed=5, first=0, stride=1, device=0, polish=0):
    from madaiemulator_b200.engine import EstimateOpts, EstimateStats
    ctx, m = ctypes.c_void_p(), ctypes.c_void_p()
    L.emub_ctx_create(device, ctypes.byref(ctx))
    X = np.zeros((4, d))
    L.mock_set_component_map(device, first, stride)
    L.emub_model_create(ctx, _P(X), d, 4, d, _P(np.zeros(4)), 1, 0, 0, ctypes.byref(m))
    L.emub_model_set_training_multi(m, _P(np.zeros((4, ncomp))), ncomp, ncomp)
    ranges = np.zeros((d + 2, 2))
    ranges[:, 0], ranges[:, 1] = -5.0, 2.0
    o = EstimateOpts()
    L.emub_estimate_default_opts(ctypes.byref(o))
    o.max_tries, o.nchains, o.seed, o.value_policy, o.first_component, o.component_stride = tries, chains, seed, policy, first, stride
    o.polish_steps = polish
    th, best = np.zeros((ncomp, d + 2)), np.zeros(ncomp)
    st = EstimateStats()
    L.mock_reset()
    rc = L.emub_estimate_thetas_multi(m, ncomp, _P(ranges), ctypes.byref(o), _P(th), _P(best), ctypes.byref(st))
    L.emub_model_destroy(m)
    return rc, th, best, st


def test_front_batches_every_chain_and_finds_the_minimum(lib):
    d, ncomp = 4, 3
    rc, th, best, st = _run(lib, d, ncomp, 0)
    assert rc == 0
    for c in range(ncomp):
        # the reference's stop rule is |g| < 0.1 (maxmultimin.c:650): the best restart ends near the minimum
        assert np.max(np.abs(th[c, 1:] - _target(c, d))) < 0.15
        assert best[c] <= 0.0 and best[c] > -0.02          # likelihood = -f
        assert abs(th[c, 0] - np.log(1.0 - best[c])) < 1e-12  # theta_0 = log sigma^2 at the optimum (maxmultimin.c:757-769)
    assert lib.mock_points() == st.evaluations and lib.mock_calls() == st.batches
    assert lib.mock_max_batch() == ncomp * 6              # all chains of all components in one batched call
    assert st.batches * 4 < st.evaluations                # ... and they stay batched
    assert st.finite_count == ncomp * 6 and st.success_count > 0


def test_value_policy_is_a_cost_decision_only(lib):
    d, ncomp = 3, 2
    res = {p: _run(lib, d, ncomp, p) for p in (0, 1, 2)}
    for p in (1, 2):
        assert np.array_equal(res[p][1], res[0][1]) and np.array_equal(res[p][2], res[0][2])
    s_ad, s_gr, s_va = res[0][3], res[1][3], res[2][3]
    assert s_va.unused_gradients == 0 and s_va.value_evaluations > 0 and s_va.repeated_points > 0
    assert s_gr.repeated_points == 0 and s_gr.value_evaluations <= 2 * 6
    assert s_va.evaluations == s_gr.evaluations + s_va.repeated_points
    assert s_gr.evaluations <= s_ad.evaluations <= s_va.evaluations
    # same seed, same answer; another seed, other restarts (but the same minimum within the stop rule)
    again = _run(lib, d, ncomp, 0)
    assert np.array_equal(again[1], res[0][1])
    other = _run(lib, d, ncomp, 0, seed=6)
    assert not np.array_equal(other[1], res[0][1]) and np.max(np.abs(other[1][:, 1:] - res[0][1][:, 1:])) < 0.3


def test_component_sharding_keeps_every_components_stream(lib):
    """component c of a sharded run (first_component / component_stride, what a rank or a device thread sets) equals
    component c of the all-in-one run bit for bit"""
    d, ncomp = 3, 4
    _, th_all, best_all, _ = _run(lib, d, ncomp, 0)
    for world in (2, 4):
        for rank in range(world):
            nloc = len(range(rank, ncomp, world))
            _, th, best, _ = _run(lib, d, nloc, 0, first=rank, stride=world, device=rank + 1)
            assert np.array_equal(th, th_all[rank::world]) and np.array_equal(best, best_all[rank::world])


def test_refinement_run_only_improves(lib):
    d, ncomp = 3, 2
    _, th0, best0, st0 = _run(lib, d, ncomp, 0)
    _, th1, best1, st1 = _run(lib, d, ncomp, 0, polish=50)
    assert np.all(best1 >= best0) and st1.evaluations > st0.evaluations
    for c in range(ncomp):
        assert np.max(np.abs(th1[c, 1:] - _target(c, d))) < 0.01  # polish_eps = 1e-3 on the gradient


def test_device_threads_shard_components_round_robin(lib):
    """emub_estimate_thetas_multi_devices: one host thread per device, component c on device c % ndev, results gathered
    back in component order and identical to the one-device run; every model is destroyed afterwards"""
    from madaiemulator_b200.engine import EstimateOpts, EstimateStats
    d, ncomp, n = 3, 5, 4
    X, Z = np.zeros((n, d)), np.zeros((n, ncomp))
    ranges = np.zeros((d + 2, 2))
    ranges[:, 0], ranges[:, 1] = -5.0, 2.0
    o = EstimateOpts()
    lib.emub_estimate_default_opts(ctypes.byref(o))
    o.max_tries, o.nchains, o.seed = 6, 6, 5
    out = {}
    for ndev in (1, 2, 3):
        devs = np.arange(10, 10 + ndev, dtype=np.int32)
        for r in range(ndev):
            lib.mock_set_component_map(10 + r, r, ndev)
        th, best, st = np.zeros((ncomp, d + 2)), np.zeros(ncomp), EstimateStats()
        lib.mock_reset()
        rc = lib.emub_estimate_thetas_multi_devices_ranges(devs.ctypes.data_as(_ip), ndev, _P(X), d, n, d, _P(Z), ncomp, ncomp, 1, 0, 0,
                                                           _P(ranges), ctypes.byref(o), _P(th), _P(best), ctypes.byref(st))
        assert rc == 0 and lib.mock_models_alive() == 0
        assert st.evaluations == lib.mock_points()
        out[ndev] = (th, best)
    for ndev in (2, 3):
        assert np.array_equal(out[ndev][0], out[1][0]) and np.array_equal(out[ndev][1], out[1][1])
    _, th_ref, best_ref, _ = _run(lib, d, ncomp, 0)
    assert np.array_equal(out[1][0], th_ref) and np.array_equal(out[1][1], best_ref)
