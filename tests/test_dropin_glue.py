"""GPU: the drop-in boundary end to end.  oracle/_ref/libemu_dropin.so is the reference's own, unmodified C
(maxWithMultiMin, doOptimizeMultiMin, modelstruct/optstruct set-up, ...) linked with integration/libemu_glue.c,
which defines evalFnMulti / gradFnMulti / evalFnGradMulti / estimateSigmaFull / alloc_emulator_struct /
emulate_point / makeCovMatrix_fnptr with the reference's signatures and forwards them to the CUDA engine.  The
same driver calls on the pure-CPU reference build (libemu_ref.so) are the expected values."""
import numpy as np
import pytest

from madaiemulator_b200 import datasets as ds
from tests.helpers import load_golden, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def oracles():
    from oracle import pyoracle as po
    if not (po.ref_available() and po.dropin_available()):
        pytest.skip("oracle/_ref drop-in build not present")
    yield po
    po.DropinOracle.reset()


@pytest.mark.parametrize("name", ["uni-simple-o1", "multi-simple-pc0-o1", "synthetic-n256-d10-o1"])
def test_reference_symbols_run_on_the_engine(oracles, name):
    po = oracles
    c = load_golden(name)
    ref = po.RefOracle(c["X"], c["y"], c["kernel"], c["order"])
    dro = po.DropinOracle(c["X"], c["y"], c["kernel"], c["order"])
    th = c["theta_less_amp"]
    # evalFnMulti / gradFnMulti through the reference's own driver
    assert relerr(dro.eval(th), ref.eval(th)) < 1e-9
    g_ref, g_dro = ref.grad(th), dro.grad(th)
    scale = np.maximum(np.abs(g_ref), 1e-3 * np.max(np.abs(g_ref)))
    assert np.max(np.abs(g_dro - g_ref) / scale) < 1e-9
    assert relerr(dro.sigma_full(th), ref.sigma_full(th)) < 1e-9
    # makeCovMatrix_fnptr
    assert relerr(dro.cov_matrix(c["theta_full"]), ref.cov_matrix(c["theta_full"]), 1e-300) < 1e-9
    # alloc_emulator_struct + emulate_point
    m1, v1 = ref.emulator(c["theta_full"]).emulate(c["pts"][:40])
    e = dro.emulator(c["theta_full"])
    m2, v2 = e.emulate(c["pts"][:40])
    assert relerr(m2, m1, 1e-3) < 1e-9
    assert np.max(np.abs(v2 - v1)) < 1e-9 * max(1.0, float(c["kappa"]))
    assert relerr(e.beta(), c["emu_beta"], 1e-6) < 1e-9
    del e


@pytest.mark.parametrize("name", ["uni-simple-o1", "multi-simple-pc0-o1"])
def test_point_list_entry_points_run_on_the_engine(oracles, name):
    """emulateAtPointList / emulateAtPoint (emulate-fns.c:73,138; the R binding's callEmulateAtList / callEmulateAtPt):
    the glue builds ONE emulator and answers the whole list with one batched prediction."""
    po = oracles
    c = load_golden(name)
    ref = po.RefOracle(c["X"], c["y"], c["kernel"], c["order"])
    dro = po.DropinOracle(c["X"], c["y"], c["kernel"], c["order"])
    pts = c["pts"][:60]
    m1, v1 = ref.emulate_at_point_list(c["theta_full"], pts)
    m2, v2 = dro.emulate_at_point_list(c["theta_full"], pts)
    assert relerr(m2, m1, 1e-3) < 1e-9
    assert np.max(np.abs(v2 - v1)) < 1e-9 * max(1.0, float(c["kappa"]))
    # emulate_model_results (emulate-fns.c:13): the same list through the resultstruct front door
    m4, v4 = ref.emulate_model_results(c["theta_full"], pts)
    m5, v5 = dro.emulate_model_results(c["theta_full"], pts)
    assert relerr(m4, m1, 1e-3) < 1e-12 and relerr(m5, m4, 1e-3) < 1e-9
    assert np.max(np.abs(v5 - v4)) < 1e-9 * max(1.0, float(c["kappa"]))
    m3, v3 = dro.emulate_at_point_list(c["theta_full"], pts[:5], single=True)
    # single points take the latency path (emub_predict_few): the same sums in another order
    assert relerr(m3, m2[:5], 1e-3) < 1e-11 and np.max(np.abs(v3 - v2[:5])) < 1e-11 * max(1.0, float(c["kappa"]))


@pytest.mark.parametrize("name", ["uni-simple-o1", "multi-simple-pc0-o1", "multi-simple-m52"])
def test_lower_level_symbols_run_on_the_engine(oracles, name):
    """makeKVector_fnptr (emulator.c:578) and chol_inverse_cov_matrix (emulate-fns.c:275) through the glue: the k-vector
    incl. the 1e-10 clamp and a query on a design point; inverse + determinant of the covariance matrix."""
    po = oracles
    c = load_golden(name)
    ref = po.RefOracle(c["X"], c["y"], c["kernel"], c["order"])
    dro = po.DropinOracle(c["X"], c["y"], c["kernel"], c["order"])
    th = c["theta_full"]
    for x in (c["pts"][0], c["pts"][3], c["X"][2], c["X"][0] + 50.0):  # far away: every entry clamps to 0
        k_ref, k_dro = ref.k_vector(th, x), dro.k_vector(th, x)
        assert relerr(k_dro, k_ref, 1e-300) < 1e-9
        assert np.array_equal(k_dro == 0.0, k_ref == 0.0)
    C = ref.cov_matrix(th)
    inv_ref, det_ref = ref.chol_inverse(C)
    inv_dro, det_dro = dro.chol_inverse(C)
    assert np.max(np.abs(inv_dro - inv_ref)) < 1e-9 * np.max(np.abs(inv_ref))
    assert np.array_equal(inv_dro, inv_dro.T)
    if det_ref > 0 and np.isfinite(det_ref):
        assert relerr(det_dro, det_ref) < 1e-9
    assert abs(np.log(det_dro) - np.linalg.slogdet(C)[1]) < 1e-9 * max(1.0, abs(np.log(det_dro))) if det_dro > 0 else True


def test_glue_recognises_a_model_by_its_contents(oracles):
    """Two models that live at the same addresses one after the other (what the R entry points do: a modelstruct per
    call) must not share an engine copy."""
    po = oracles
    c = load_golden("uni-simple-o1")
    th = c["theta_less_amp"]
    vals = []
    for scale in (1.0, 3.0, 1.0):
        dro = po.DropinOracle(c["X"], scale * c["y"], c["kernel"], c["order"])
        ref = po.RefOracle(c["X"], scale * c["y"], c["kernel"], c["order"])
        assert relerr(dro.sigma_full(th), ref.sigma_full(th)) < 1e-9
        vals.append(dro.sigma_full(th))
        del dro, ref
    assert relerr(vals[1], 9.0 * vals[0]) < 1e-9 and vals[2] == vals[0]


def test_reference_restart_driver_on_the_engine(oracles):
    """The reference's unmodified maxWithMultiMin / doOptimizeMultiMin, every likelihood call served by the GPU:
    same seed, same start points -> the same optimum as the all-CPU reference."""
    po = oracles
    c = load_golden("uni-simple-o1")
    ref = po.RefOracle(c["X"], c["y"], c["kernel"], c["order"])
    dro = po.DropinOracle(c["X"], c["y"], c["kernel"], c["order"])
    best_ref, th_ref = ref.max_with_multimin(6, 5)
    best_dro, th_dro = dro.max_with_multimin(6, 5)
    assert abs(best_dro - best_ref) < 1e-6 * max(1.0, abs(best_ref))
    assert np.max(np.abs(th_dro - th_ref)) < 1e-4


def test_emuplusplus_front_end_on_the_engine():
    """The reference's C++ wrapper (src/EmuPlusPlus.cpp) and its example driver (test/emuplusplus-test/src/example.cpp),
    both unmodified, linked against the drop-in library: same printed means / errors as on the CPU reference build."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref_bin = os.path.join(root, "oracle", "_ref", "emuplusplus_ref")
    dro_bin = os.path.join(root, "oracle", "_ref", "emuplusplus_dropin")
    if not (os.path.exists(ref_bin) and os.path.exists(dro_bin)):
        pytest.skip("oracle/_ref EmuPlusPlus builds not present")
    snap = os.path.join(root, "tests", "golden", "cli", "multi-simple-o0.snapshot")
    pts = "".join(open(os.path.join(root, "tests", "golden", "cli", "multi-simple.points")).readlines()[:25]).encode()
    a = subprocess.run([ref_bin, snap], input=pts, capture_output=True, check=True, timeout=300).stdout.decode().split("\n")
    b = subprocess.run([dro_bin, snap], input=pts, capture_output=True, check=True, timeout=300).stdout.decode().split("\n")
    assert len(a) == len(b) and len(a) > 70
    for la, lb in zip(a, b):
        if la.startswith("# mean:") or la.startswith("# err:"):
            va = np.array(la.split(":")[1].split(), dtype=np.float64)
            vb = np.array(lb.split(":")[1].split(), dtype=np.float64)
            assert np.allclose(va, vb, rtol=2e-6, atol=1e-9)  # cout prints 6 significant digits
        else:
            assert la == lb


def test_reference_cli_trains_on_the_engine(tmp_path):
    """BASELINE configs 0/1 end to end through the reference's UNCHANGED CLI: `interactive_emulator estimate_thetas`
    (input reader, PCA, model structs and snapshot writer are the reference's; estimate_thetas_threaded is the glue
    -> batched GPU restarts).  north_star: the likelihood reached must be no worse than the reference's own training
    (tests/golden/cli/*.snapshot were trained by the reference CLI on the CPU with 8 x 50 restarts)."""
    import ctypes
    import os
    import subprocess
    from madaiemulator_b200 import engine
    from oracle.pyoracle import PortOracle
    from tests.test_interactive_stream import _Snap
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cli = os.path.join(root, "oracle", "_ref", "interactive_emulator_dropin")
    if not os.path.exists(cli):
        pytest.skip("oracle/_ref/interactive_emulator_dropin not built")
    H = engine.host_lib()
    H.emub_snapshot_load_path.restype = ctypes.POINTER(_Snap)
    H.emub_snapshot_load_path.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]

    def thetas_of(path):
        err = ctypes.create_string_buffer(256)
        sp = H.emub_snapshot_load_path(path.encode(), err, 256)
        assert sp, err.value
        s = sp.contents
        out = []
        for c in range(s.nr):
            comp = s.components[c]
            n, d = comp.nmodel_points, comp.nparams
            out.append((np.ctypeslib.as_array(comp.xmodel, (n * d,)).reshape(n, d).copy(),
                        np.ctypeslib.as_array(comp.training_vector, (n,)).copy(),
                        np.ctypeslib.as_array(comp.thetas, (comp.nthetas,)).copy(), comp.regression_order))
        return out

    # the reference's input file, reconstructed from its own snapshot (X and the training matrix are stored verbatim)
    ref_snap = os.path.join(root, "tests", "golden", "cli", "uni-simple-o1.snapshot")
    tok = open(ref_snap).read().split()
    nt, nr, d, n = int(tok[0]), int(tok[1]), int(tok[2]), int(tok[3])
    inp = tmp_path / "input_model_file.dat"
    inp.write_text("%d\n%d\n%d\n" % (nt, d, n) + "\n".join(tok[6:6 + n * d]) + "\n" + "\n".join(tok[6 + n * d:6 + n * d + n * nt]) + "\n")
    out_snap = tmp_path / "trained_on_gpu.snapshot"
    env = dict(os.environ, EMUB_TRIES="400", EMUB_SLOTS="64", EMUB_SEED="5")  # the reference run: 8 threads x 50 restarts
    subprocess.run([cli, "estimate_thetas", str(inp), str(out_snap), "--regression_order=1"], check=True, timeout=600, env=env,
                   stdout=subprocess.DEVNULL)
    ours, ref = thetas_of(str(out_snap)), thetas_of(ref_snap)
    assert len(ours) == len(ref) == 1
    for (X1, y1, th1, o1), (X0, y0, th0, o0) in zip(ours, ref):
        assert np.allclose(X1, X0) and np.allclose(y1, y0, atol=1e-12) and o1 == o0 == 1
        po = PortOracle(X0, y0, 1, o0)
        l_ours = -po.loglik_grad(th1[1:], want_grad=False)["negL"]
        l_ref = -po.loglik_grad(th0[1:], want_grad=False)["negL"]
        assert l_ours >= l_ref - 1e-3 * max(1.0, abs(l_ref)), (l_ours, l_ref)


# ---- the multivariate front doors over the batched entry points (integration/multivar_glue.c) -------------
def _multi_bin(name):
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "oracle", "_ref", name)
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/%s not built" % name)
    return root, path


@pytest.mark.parametrize("flag,golden,nr", [(None, "multi-simple-o0.interactive.txt", None),
                                            ("--pca_output", "multi-simple-o0.interactive_pca.txt", 5)])
def test_reference_cli_interactive_mode_with_batched_front_doors(flag, golden, nr):
    """The reference's unmodified interactive_emulator.c; alloc_multi_emulator / emulate_point_multi[_pca] are the
    glue's: one engine model with all PCA components, one device pass (back-projection included) per point."""
    import os
    import subprocess
    from tests.test_interactive_stream import _compare_protocol
    root, cli = _multi_bin("interactive_emulator_dropin_multi")
    cdir = os.path.join(root, "tests", "golden", "cli")
    inp = open(os.path.join(cdir, "multi-simple.points"), "rb").read()
    cmd = [cli, "interactive_mode", os.path.join(cdir, "multi-simple-o0.snapshot")] + ([flag] if flag else [])
    out = subprocess.run(cmd, input=inp, capture_output=True, check=True, timeout=300).stdout.decode()
    _compare_protocol(out, os.path.join(cdir, golden), 6, 0 if flag else 1 + 3 + 1 + 12, nr)


def test_reference_cli_single_output_model_with_batched_front_doors():
    """a one-output model (nt = nr = 1, BASELINE config 0): the few-points path without the component table"""
    import os
    import subprocess
    from tests.test_interactive_stream import _compare_protocol
    root, cli = _multi_bin("interactive_emulator_dropin_multi")
    cdir = os.path.join(root, "tests", "golden", "cli")
    inp = open(os.path.join(cdir, "uni-simple.points"), "rb").read()
    out = subprocess.run([cli, "interactive_mode", os.path.join(cdir, "uni-simple-o1.snapshot")], input=inp, capture_output=True,
                         check=True, timeout=300).stdout.decode()
    _compare_protocol(out, os.path.join(cdir, "uni-simple-o1.interactive.txt"), 1, 1 + 1 + 1 + 2)


def test_emuplusplus_with_batched_front_doors():
    import os
    import subprocess
    root, dro_bin = _multi_bin("emuplusplus_dropin_multi")
    ref_bin = os.path.join(root, "oracle", "_ref", "emuplusplus_ref")
    if not os.path.exists(ref_bin):
        pytest.skip("oracle/_ref/emuplusplus_ref not built")
    snap = os.path.join(root, "tests", "golden", "cli", "multi-simple-o0.snapshot")
    pts = "".join(open(os.path.join(root, "tests", "golden", "cli", "multi-simple.points")).readlines()[:25]).encode()
    a = subprocess.run([ref_bin, snap], input=pts, capture_output=True, check=True, timeout=300).stdout.decode().split("\n")
    b = subprocess.run([dro_bin, snap], input=pts, capture_output=True, check=True, timeout=300).stdout.decode().split("\n")
    assert len(a) == len(b) and len(a) > 70
    for la, lb in zip(a, b):
        if la.startswith("# mean:") or la.startswith("# err:"):
            va = np.array(la.split(":")[1].split(), dtype=np.float64)
            vb = np.array(lb.split(":")[1].split(), dtype=np.float64)
            assert np.allclose(va, vb, rtol=2e-6, atol=1e-9)
        else:
            assert la == lb


def test_reference_cli_trains_all_components_in_one_front(tmp_path):
    """BASELINE config 1 (multi-simple: 3 parameters, 6 outputs -> 5 PCA components) trained through the reference's
    UNCHANGED CLI with estimate_multi bound to the batched driver: the restart chains of all 5 components share one
    evaluation front.  Every component must reach a likelihood no worse than the reference's own CPU training."""
    import ctypes
    import os
    import subprocess
    from madaiemulator_b200 import engine
    from oracle.pyoracle import PortOracle
    from tests.test_interactive_stream import _Snap
    root, cli = _multi_bin("interactive_emulator_dropin_multi")
    H = engine.host_lib()
    H.emub_snapshot_load_path.restype = ctypes.POINTER(_Snap)
    H.emub_snapshot_load_path.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]

    def components_of(path):
        err = ctypes.create_string_buffer(256)
        sp = H.emub_snapshot_load_path(path.encode(), err, 256)
        assert sp, err.value
        s = sp.contents
        out = []
        for c in range(s.nr):
            comp = s.components[c]
            n, d = comp.nmodel_points, comp.nparams
            out.append((np.ctypeslib.as_array(comp.xmodel, (n * d,)).reshape(n, d).copy(),
                        np.ctypeslib.as_array(comp.training_vector, (n,)).copy(),
                        np.ctypeslib.as_array(comp.thetas, (comp.nthetas,)).copy(), comp.regression_order))
        return out

    ref_snap = os.path.join(root, "tests", "golden", "cli", "multi-simple-o0.snapshot")
    tok = open(ref_snap).read().split()
    nt, nr, d, n = int(tok[0]), int(tok[1]), int(tok[2]), int(tok[3])
    inp = tmp_path / "multi-test-input.dat"
    inp.write_text("%d\n%d\n%d\n" % (nt, d, n) + "\n".join(tok[6:6 + n * d]) + "\n" + "\n".join(tok[6 + n * d:6 + n * d + n * nt]) + "\n")
    out_snap = tmp_path / "trained_on_gpu.snapshot"
    env = dict(os.environ, EMUB_TRIES="400", EMUB_SLOTS="16", EMUB_SEED="7")
    subprocess.run([cli, "estimate_thetas", str(inp), str(out_snap), "--regression_order=0"], check=True, timeout=900, env=env,
                   stdout=subprocess.DEVNULL)
    ours, ref = components_of(str(out_snap)), components_of(ref_snap)
    assert len(ours) == len(ref) == nr == 5
    for (X1, y1, th1, o1), (X0, y0, th0, o0) in zip(ours, ref):
        # the PCA is the reference's own code in both runs: same components up to rounding
        assert np.allclose(X1, X0) and np.allclose(y1, y0, atol=1e-9) and o1 == o0 == 0
        po = PortOracle(X0, y0, 1, o0)
        l_ours = -po.loglik_grad(th1[1:], want_grad=False)["negL"]
        l_ref = -po.loglik_grad(th0[1:], want_grad=False)["negL"]
        assert l_ours >= l_ref - 1e-3 * max(1.0, abs(l_ref)), (l_ours, l_ref)


def test_batched_front_doors_shard_over_two_gpus(tmp_path):
    """EMUB_DEVICES=0,1: estimate_multi sends component c to device c mod 2 (no exchange between devices) and
    alloc_multi_emulator replicates the factors; same seed -> byte-identical snapshot and answers as on one GPU."""
    import os
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root, cli = _multi_bin("interactive_emulator_dropin_multi")
    cdir = os.path.join(root, "tests", "golden", "cli")
    tok = open(os.path.join(cdir, "multi-simple-o0.snapshot")).read().split()
    nt, d, n = int(tok[0]), int(tok[2]), int(tok[3])
    inp = tmp_path / "multi-test-input.dat"
    inp.write_text("%d\n%d\n%d\n" % (nt, d, n) + "\n".join(tok[6:6 + n * d]) + "\n" + "\n".join(tok[6 + n * d:6 + n * d + n * nt]) + "\n")
    outs = []
    for devs in ("0", "0,1"):
        env = dict(os.environ, EMUB_TRIES="48", EMUB_SLOTS="16", EMUB_SEED="3", EMUB_DEVICES=devs)
        snap = tmp_path / ("trained_%d.snapshot" % len(outs))
        subprocess.run([cli, "estimate_thetas", str(inp), str(snap), "--regression_order=0"], check=True, timeout=900, env=env,
                       stdout=subprocess.DEVNULL)
        pts = open(os.path.join(cdir, "multi-simple.points"), "rb").read()
        ans = subprocess.run([cli, "interactive_mode", str(snap), "--quiet"], input=pts, capture_output=True, check=True, timeout=300,
                             env=env).stdout
        outs.append((snap.read_bytes(), ans))
    assert outs[0][0] == outs[1][0]
    assert outs[0][1] == outs[1][1]


def test_glue_model_cache_is_bounded(oracles):
    """Every engine model keeps EMUB_SLOTS factorisation slots on the device; a caller that walks over many
    modelstructs (the per-component loop of estimate_multi, the R entry points) must not accumulate them: the glue
    keeps at most EMUB_GLUE_MODELS (default 8), least recently used first out, and a revisited model is rebuilt."""
    po = oracles
    po.DropinOracle.reset()
    L = po.DropinOracle.lib()
    c = load_golden("uni-simple-o1")
    th = c["theta_less_amp"]
    models = [po.DropinOracle(c["X"], (1.0 + 0.1 * k) * c["y"], c["kernel"], c["order"]) for k in range(12)]
    first = [m.sigma_full(th) for m in models]
    assert L.libemu_glue_model_count() <= 8
    again = [m.sigma_full(th) for m in models]
    assert first == again
    ref = po.RefOracle(c["X"], 1.3 * c["y"], c["kernel"], c["order"])
    assert relerr(first[3], ref.sigma_full(th)) < 1e-9
    # a model with a live emulator_struct is pinned while the others come and go
    e = models[0].emulator(c["theta_full"])
    m0, v0 = e.emulate(c["pts"][:5])
    for m in models[1:]:
        m.sigma_full(th)
    m1, v1 = e.emulate(c["pts"][:5])
    assert np.array_equal(m0, m1) and np.array_equal(v0, v1)
    del e, models
    po.DropinOracle.reset()
