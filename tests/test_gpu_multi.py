"""GPU: the component-batched callers: estimate_multi / alloc_multi_emulator / emulate_point_multi
re-expressed over the batched C-ABI -- all PCA components of a multivariate model share one design, one evaluation
front and one prediction pass, with the back-projection on the device."""
import os

import numpy as np
import pytest

from madaiemulator_b200 import datasets as ds
from tests.helpers import relerr

pytestmark = pytest.mark.gpu
TOL = 1e-9
MULTI_SIMPLE = "/root/reference/test/multi-simple/multi-test-input.dat"


@pytest.fixture(scope="module")
def ctx():
    from madaiemulator_b200 import engine
    c = engine.Context(0)
    yield c
    c.close()


def _multi_problem(n=150, d=3, nt=6):
    X, Y = ds.synthetic_model(n, d, nt)
    pca = ds.pca_decompose(Y, 0.99)
    return X, Y, pca


def test_component_batch_equals_single_component_models(ctx):
    from madaiemulator_b200 import engine
    X, Y, pca = _multi_problem()
    Z, nr = pca["Z"], pca["nr"]
    assert nr >= 2
    m = engine.Model(ctx, X, Z[:, 0], 1, 1, max_slots=8)
    m.set_training_multi(Z)
    rng = np.random.default_rng(0)
    B = 11
    ths = np.stack([np.concatenate([[rng.uniform(-5, -2)], rng.uniform(0.0, 1.5, X.shape[1])]) for _ in range(B)])
    comp = rng.integers(0, nr, B)
    r = m.loglik_grad_batch(ths, comp=comp)
    for c in range(nr):
        single = engine.Model(ctx, X, Z[:, c], 1, 1, max_slots=8)
        idx = np.where(comp == c)[0]
        if len(idx):
            rs = single.loglik_grad_batch(ths[idx])
            assert np.array_equal(rs["negL"], r["negL"][idx]) and np.array_equal(rs["grad"], r["grad"][idx])
            assert np.array_equal(rs["sigma2"], r["sigma2"][idx])
        single.close()
    # and against the CPU oracle
    from oracle.pyoracle import PortOracle
    for b in (0, B - 1):
        ref = PortOracle(X, Z[:, comp[b]], 1, 1).loglik_grad(ths[b])
        assert relerr(r["negL"][b], ref["negL"]) < TOL
    m.close()


def test_predict_multi_matches_oracle_backprojection(ctx):
    """emulate_point_multi: per-component emulators + back-projection (multivar_support.c:126-151)."""
    from madaiemulator_b200 import engine
    from oracle.pyoracle import PortOracle, backproject
    X, Y, pca = _multi_problem()
    Z, nr, nt = pca["Z"], pca["nr"], Y.shape[1]
    d = X.shape[1]
    m = engine.Model(ctx, X, Z[:, 0], 1, 1, max_slots=4)
    m.set_training_multi(Z)
    rng = np.random.default_rng(1)
    thetas = np.stack([np.concatenate([[rng.uniform(-1, 0.5), rng.uniform(-5, -3)], rng.uniform(0.3, 1.2, d)]) for _ in range(nr)])
    emus = [m.emulator(thetas[c], comp=c) for c in range(nr)]
    pts = ds.synthetic_queries(300, d)
    pts[0] = X[7]
    mean, var = engine.predict_multi(emus, pts, pca["mean"], pca["evecs"], pca["evals"])
    mp, vp = engine.predict_multi(emus, pts)
    assert mean.shape == (300, nt) and mp.shape == (300, nr)
    om = np.empty((300, nr))
    ov = np.empty((300, nr))
    for c in range(nr):
        om[:, c], ov[:, c] = PortOracle(X, Z[:, c], 1, 1).emulator(thetas[c]).emulate(pts)
    assert relerr(mp, om, 1e-3) < TOL
    assert np.max(np.abs(vp - ov)) < TOL * 2.0
    for q in (0, 1, 150, 299):
        mo, vo = backproject(pca["mean"], pca["evecs"], pca["evals"], om[q], ov[q])
        assert np.max(np.abs(mean[q] - mo)) < TOL * max(1.0, np.max(np.abs(mo)))
        assert np.max(np.abs(var[q] - vo)) < TOL * max(1.0, np.max(np.abs(vo)))
    for e in emus:
        e.close()
    m.close()


def test_predict_multi_many_observables(ctx):
    """more observables than the back-projection buffers start out with (64): they grow on demand; the projection of
    3 components onto 200 observables equals the host formula, and a later narrow call still works"""
    from madaiemulator_b200 import engine
    n, d, nr, nt = 200, 3, 3, 200
    X = ds.synthetic_design(n, d)
    rng = np.random.default_rng(4)
    Z = rng.normal(size=(n, nr))
    m = engine.Model(ctx, X, Z[:, 0], 1, 0, max_slots=2)
    m.set_training_multi(Z)
    thetas = np.stack([np.concatenate([[0.0, -3.0], rng.uniform(0.3, 1.0, d)]) for _ in range(nr)])
    emus = [m.emulator(thetas[c], comp=c) for c in range(nr)]
    pts = ds.synthetic_queries(500, d)
    ybar, U, lam = rng.normal(size=nt), rng.normal(size=(nt, nr)), rng.uniform(0.5, 2.0, nr)
    mp, vp = engine.predict_multi(emus, pts)
    mean, var = engine.predict_multi(emus, pts, ybar, U, lam)
    assert mean.shape == (500, nt)
    em = ybar[None, :] + (mp * np.sqrt(lam)[None, :]) @ U.T
    ev = (vp * lam[None, :]) @ (U * U).T
    assert np.max(np.abs(mean - em)) < 1e-12 * max(1.0, np.max(np.abs(em)))
    assert np.max(np.abs(var - ev)) < 1e-12 * max(1.0, np.max(np.abs(ev)))
    mp2, vp2 = engine.predict_multi(emus, pts)
    assert np.array_equal(mp, mp2) and np.array_equal(vp, vp2)
    # the latency path for a handful of points: same answers to rounding, in both output spaces
    mf, vf = engine.predict_multi(emus, pts[:3], ybar, U, lam, few=True)
    assert np.max(np.abs(mf - mean[:3])) < 1e-11 * max(1.0, np.max(np.abs(mean))) and np.max(np.abs(vf - var[:3])) < 1e-11 * max(1.0, np.max(np.abs(var)))
    mf, vf = engine.predict_multi(emus, pts[:1], few=True)
    assert np.max(np.abs(mf - mp[:1])) < 1e-11 and np.max(np.abs(vf - vp[:1])) < 1e-11
    # emulators rebuilt with other length scales (same amplitude and nugget, very likely the same device addresses):
    # the cached table of the few-points path must not survive them
    for e in emus:
        e.close()
    thetas2 = thetas.copy()
    thetas2[:, 2:] += 0.4
    emus = [m.emulator(thetas2[c], comp=c) for c in range(nr)]
    mb2, vb2 = engine.predict_multi(emus, pts[:2])
    mf2, vf2 = engine.predict_multi(emus, pts[:2], few=True)
    assert np.max(np.abs(mf2 - mb2)) < 1e-11 and np.max(np.abs(vf2 - vb2)) < 1e-11
    assert np.max(np.abs(mb2 - mp[:2])) > 1e-6  # and they really are different emulators
    with pytest.raises(engine.EmubError):
        engine.predict_multi(emus, pts, np.zeros(2000), np.zeros((2000, nr)), lam)
    for e in emus:
        e.close()
    m.close()


def test_estimate_multi_merges_fronts(ctx):
    """All components' restart chains in one evaluation front; component k's result equals a single-component run."""
    from madaiemulator_b200 import engine
    X, Y, pca = _multi_problem(n=100)
    Z, nr = pca["Z"], pca["nr"]
    m = engine.Model(ctx, X, Z[:, 0], 1, 0, max_slots=32)
    m.set_training_multi(Z)
    th, best, st = engine.estimate_thetas_multi(m, nr, max_tries=6, nchains=6, seed=3)
    assert st["rc"] == 0 and np.all(np.isfinite(best))
    assert st["evaluations"] / st["batches"] > 6  # wider than any single component's front
    k = nr - 1
    single = engine.Model(ctx, X, Z[:, k], 1, 0, max_slots=8)
    th1, best1, _ = engine.estimate_thetas(single, max_tries=6, nchains=6, seed=(3 + 0x9E3779B97F4A7C15 * k) % (1 << 64))
    assert np.array_equal(th1, th[k]) and best1 == best[k]
    single.close()
    m.close()


@pytest.mark.skipif(not os.path.exists(MULTI_SIMPLE), reason="reference fixture not on this machine")
def test_multi_simple_fixture_shape():
    X, Y = ds.load_input_model_file(MULTI_SIMPLE)
    assert X.shape == (100, 3) and Y.shape == (100, 6)
    assert ds.pca_decompose(Y, 0.99)["nr"] <= 5


def test_components_shard_over_two_gpus(ctx):
    """cfg4 pattern: PCA components sharded over the GPUs of one box, thetas gathered on the host -- identical to
    the single-device run (same per-component start-point streams)."""
    import torch
    from madaiemulator_b200 import engine
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    X, Y, pca = _multi_problem(n=100)
    Z, nr = pca["Z"], pca["nr"]
    m = engine.Model(ctx, X, Z[:, 0], 1, 0, max_slots=32)
    m.set_training_multi(Z)
    th1, best1, _ = engine.estimate_thetas_multi(m, nr, max_tries=6, nchains=6, seed=3)
    m.close()
    th2, best2, st = engine.estimate_thetas_multi_devices([0, 1], X, Z, 1, 0, max_tries=6, nchains=6, seed=3, max_slots=32)
    assert st["rc"] == 0
    assert np.array_equal(th1, th2) and np.array_equal(best1, best2)


@pytest.mark.gpu
def test_uploads_are_ordered_before_the_kernels_that_read_them():
    """Regression: the training vector used to go up with a plain cudaMemcpy (pageable source, <= 64 KB: staged, the DMA
    runs on the legacy default stream) while k_build_yh read it on a non-blocking stream -- now and then a model
    trained on whatever the recycled device buffer held, i.e. the PREVIOUS model's training vector.  Models are created
    on recycled memory with alternating data and evaluated at once; every one must see its own data."""
    from madaiemulator_b200 import engine
    n, d = 300, 3
    X = ds.synthetic_design(n, d)
    ys = [ds.synthetic_response(X, t) for t in range(3)]
    th = ds.default_theta_less_amp(d)
    ctx = engine.Context(0)
    expect = []
    for y in ys:
        m = engine.Model(ctx, X, y, 1, 0, max_slots=2)
        expect.append(m.loglik_grad_batch(th[None, :])["negL"][0])
        m.close()
    assert len(set(expect)) == 3
    for i in range(60):
        k = (i * 7) % 3
        m = engine.Model(ctx, X, ys[(k + 1) % 3], 1, 0, max_slots=2)
        if i % 2:
            m.set_training_multi(np.stack([ys[k], ys[(k + 2) % 3]], axis=1))
            got = m.loglik_grad_batch(np.stack([th, th]), want_grad=False, comp=[0, 1])["negL"]
            assert got[0] == expect[k] and got[1] == expect[(k + 2) % 3], i
        else:
            m.set_training(ys[k])
            assert m.loglik_grad_batch(th[None, :], want_grad=False)["negL"][0] == expect[k], i
        m.close()
    ctx.close()


@pytest.mark.gpu
def test_component_sharding_does_not_change_a_bit():
    """What bench.py's strong-scaling section asserts at N > 1 (`sharded_identical`), on one device: all PCA components
    in one evaluation front against one component at a time with first_component / component_stride (what every rank
    does when the components are shared out) -- same thetas, same likelihoods, bit for bit."""
    from madaiemulator_b200 import engine
    n, d, ncomp, restarts = 700, 5, 4, 4
    X, Y = ds.synthetic_model(n, d, nt=ncomp + 1)
    Z = np.ascontiguousarray(ds.pca_decompose(Y, vfrac=2.0)["Z"][:, :ncomp])
    ranges = engine.optimization_ranges(engine.POWEREXP, X)
    ctx = engine.Context(0)

    def train(components, first, stride):
        m = engine.Model(ctx, X, Z[:, components[0]], engine.POWEREXP, 0, max_slots=restarts * len(components))
        m.set_training_multi(Z[:, components])
        th, best, st = engine.estimate_thetas_multi(m, len(components), ranges, max_tries=restarts, nchains=restarts, seed=3,
                                                    step_max=6, first_component=first, component_stride=stride)
        m.close()
        return th, best, st

    th_all, best_all, st_all = train(list(range(ncomp)), 0, 1)
    assert st_all["value_evaluations"] > 0
    for rep in range(2):
        for c in range(ncomp):
            th, best, _ = train([c], c, ncomp)
            assert np.array_equal(th[0], th_all[c]) and best[0] == best_all[c], (rep, c)
    # two components per "rank"
    for r in range(2):
        comps = [r, r + 2]
        th, best, _ = train(comps, r, 2)
        assert np.array_equal(th, th_all[comps]) and np.array_equal(best, best_all[comps])
    ctx.close()
