"""MODEL_SNAPSHOT_FILE reader and the streaming interactive_mode, against fixtures produced by
the reference's own CLI (tests/golden/cli/README.md)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI_DIR = os.path.join(ROOT, "tests", "golden", "cli")
TOOL = os.path.join(ROOT, "madaiemulator_b200", "host", "emub_interactive_emulator")
DROPIN_CLI = os.path.join(ROOT, "oracle", "_ref", "interactive_emulator_dropin")

_dp = ctypes.POINTER(ctypes.c_double)


class _Comp(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int) for k in ("nthetas", "nparams", "nmodel_points", "nemulate_points", "regression_order",
                                            "nregression_fns", "fixed_nugget_mode", "cov_fn_index", "use_data_scales")] + \
               [("fixed_nugget", ctypes.c_double)] + [(k, _dp) for k in ("grad_ranges", "xmodel", "training_vector", "thetas", "sample_scales")]


class _Snap(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int) for k in ("nt", "nr", "nparams", "nmodel_points", "cov_fn_index", "regression_order")] + \
               [(k, _dp) for k in ("xmodel", "training_matrix", "training_mean", "pca_evals_r", "pca_evecs_r", "pca_zmatrix")] + \
               [("components", ctypes.POINTER(_Comp))]


def _tokens(path):
    return open(path).read().split()


@pytest.mark.parametrize("name,nt,nr,d,n,order", [("uni-simple-o1", 1, 1, 1, 34, 1), ("multi-simple-o0", 6, 5, 3, 100, 0)])
def test_snapshot_reader(name, nt, nr, d, n, order):
    from madaiemulator_b200 import engine
    H = engine.host_lib()
    H.emub_snapshot_load_path.restype = ctypes.POINTER(_Snap)
    H.emub_snapshot_load_path.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
    H.emub_snapshot_free.argtypes = [ctypes.POINTER(_Snap)]
    H.emub_snapshot_free.restype = None
    path = os.path.join(CLI_DIR, name + ".snapshot")
    err = ctypes.create_string_buffer(256)
    sp = H.emub_snapshot_load_path(path.encode(), err, 256)
    assert sp, err.value
    s = sp.contents
    assert (s.nt, s.nr, s.nparams, s.nmodel_points, s.cov_fn_index, s.regression_order) == (nt, nr, d, n, 1, order)
    tok = _tokens(path)
    X = np.array(tok[6:6 + n * d], dtype=np.float64)
    assert np.array_equal(np.ctypeslib.as_array(s.xmodel, (n * d,)), X)
    Y = np.array(tok[6 + n * d:6 + n * d + n * nt], dtype=np.float64).reshape(n, nt)
    assert np.array_equal(np.ctypeslib.as_array(s.training_matrix, (n * nt,)), Y.ravel())
    mean = np.ctypeslib.as_array(s.training_mean, (nt,))
    assert np.allclose(mean, Y.mean(axis=0), rtol=1e-14, atol=1e-15)
    # the last component block ends the file: its sample scales are the last d tokens, its thetas the nthetas before
    last = s.components[nr - 1]
    assert last.nthetas == d + 2 and last.regression_order == order and last.nregression_fns == 1 + order * d
    assert np.array_equal(np.ctypeslib.as_array(last.sample_scales, (d,)), np.array(tok[-d:], dtype=np.float64))
    assert np.array_equal(np.ctypeslib.as_array(last.thetas, (d + 2,)), np.array(tok[-d - (d + 2):-d], dtype=np.float64))
    # every block repeats the design, and its training vector is the matching z-matrix column
    Z = np.ctypeslib.as_array(s.pca_zmatrix, (n * nr,)).reshape(n, nr)
    for c in range(nr):
        comp = s.components[c]
        assert np.array_equal(np.ctypeslib.as_array(comp.xmodel, (n * d,)), X)
        assert np.array_equal(np.ctypeslib.as_array(comp.training_vector, (n,)), Z[:, c])
    H.emub_snapshot_free(sp)
    # malformed input is reported, not crashed on
    bad = os.path.join("/tmp", "bad_%s.snapshot" % name)
    open(bad, "w").write(" ".join(tok[:50]))
    assert not H.emub_snapshot_load_path(bad.encode(), err, 256)
    assert b"malformed" in err.value


@pytest.mark.parametrize("name", ["uni-simple-o1", "multi-simple-o0"])
def test_snapshot_writer_is_byte_identical_to_the_reference_dump(name, tmp_path):
    """load -> save reproduces the file the reference CLI wrote (dump_multi_modelstruct, multi_modelstruct.c:346-401)."""
    from madaiemulator_b200 import engine
    H = engine.host_lib()
    H.emub_snapshot_load_path.restype = ctypes.POINTER(_Snap)
    H.emub_snapshot_load_path.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
    H.emub_snapshot_save_path.argtypes = [ctypes.POINTER(_Snap), ctypes.c_char_p]
    H.emub_snapshot_free.argtypes = [ctypes.POINTER(_Snap)]
    H.emub_snapshot_free.restype = None
    path = os.path.join(CLI_DIR, name + ".snapshot")
    err = ctypes.create_string_buffer(256)
    sp = H.emub_snapshot_load_path(path.encode(), err, 256)
    assert sp, err.value
    out = str(tmp_path / "copy.snapshot")
    assert H.emub_snapshot_save_path(sp, out.encode()) == 0
    H.emub_snapshot_free(sp)
    assert open(out, "rb").read() == open(path, "rb").read()


def test_snapshot_from_arrays_round_trip(tmp_path):
    from madaiemulator_b200 import engine
    H = engine.host_lib()
    H.emub_snapshot_from_arrays.restype = ctypes.POINTER(_Snap)
    H.emub_snapshot_from_arrays.argtypes = [_dp, ctypes.c_int, ctypes.c_int, _dp, ctypes.c_int, _dp, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    H.emub_snapshot_load_path.restype = ctypes.POINTER(_Snap)
    H.emub_snapshot_load_path.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
    H.emub_snapshot_save_path.argtypes = [ctypes.POINTER(_Snap), ctypes.c_char_p]
    H.emub_snapshot_free.argtypes = [ctypes.POINTER(_Snap)]
    H.emub_snapshot_free.restype = None
    rng = np.random.default_rng(3)
    n, d, nr = 40, 3, 2
    X = np.ascontiguousarray(rng.uniform(-1, 1, (n, d)))
    Z = np.ascontiguousarray(rng.normal(size=(n, nr)))
    th = np.ascontiguousarray(rng.uniform(-3, 1, (nr, d + 2)))
    sp = H.emub_snapshot_from_arrays(X.ctypes.data_as(_dp), n, d, Z.ctypes.data_as(_dp), nr, th.ctypes.data_as(_dp), d + 2, 1, 1)
    assert sp
    out = str(tmp_path / "made.snapshot")
    assert H.emub_snapshot_save_path(sp, out.encode()) == 0
    H.emub_snapshot_free(sp)
    err = ctypes.create_string_buffer(256)
    sp = H.emub_snapshot_load_path(out.encode(), err, 256)
    assert sp, err.value
    s = sp.contents
    assert (s.nt, s.nr, s.nparams, s.nmodel_points, s.cov_fn_index, s.regression_order) == (nr, nr, d, n, 1, 1)
    # "%.17lf" keeps 17 decimals, not 17 significant digits: values are reproduced to 1e-17 absolute
    assert np.max(np.abs(np.ctypeslib.as_array(s.xmodel, (n * d,)) - X.ravel())) < 1e-16
    mean = np.ctypeslib.as_array(s.training_mean, (nr,))
    assert np.allclose(mean, Z.mean(axis=0), atol=1e-15)
    for c in range(nr):
        comp = s.components[c]
        assert comp.nregression_fns == 1 + d and comp.nthetas == d + 2
        assert np.max(np.abs(np.ctypeslib.as_array(comp.training_vector, (n,)) - (Z[:, c] - Z[:, c].mean()))) < 1e-15
        assert np.max(np.abs(np.ctypeslib.as_array(comp.thetas, (d + 2,)) - th[c])) < 1e-16
    # identity back-projection
    assert np.array_equal(np.ctypeslib.as_array(s.pca_evecs_r, (nr * nr,)).reshape(nr, nr), np.eye(nr))
    assert np.array_equal(np.ctypeslib.as_array(s.pca_evals_r, (nr,)), np.ones(nr))
    H.emub_snapshot_free(sp)


def _parse(H, text, maxvals, threads):
    buf = ctypes.create_string_buffer(text, len(text) + 1)
    out = np.full(maxvals + 4, -777.0)
    consumed = ctypes.c_size_t(0)
    bad = ctypes.c_int(0)
    k = H.emub_parse_doubles(buf, len(text), out.ctypes.data_as(_dp), maxvals, threads, ctypes.byref(consumed), ctypes.byref(bad))
    assert np.all(out[maxvals:] == -777.0)  # never writes past max
    assert buf.raw[:len(text)] == text       # the text is left as it was
    return out[:k], consumed.value, bad.value


def test_parallel_text_parser():
    """the input side of the stream: same values as strtod/float() token by token, whatever the thread count"""
    from madaiemulator_b200 import engine
    H = engine.host_lib()
    H.emub_parse_doubles.restype = ctypes.c_size_t
    H.emub_parse_doubles.argtypes = [ctypes.c_char_p, ctypes.c_size_t, _dp, ctypes.c_size_t, ctypes.c_int,
                                     ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_int)]
    rng = np.random.default_rng(11)
    vals = np.concatenate([rng.uniform(-3, 3, 30000), rng.normal(size=2000) * 1e-300, rng.normal(size=2000) * 1e300,
                           [0.0, -0.0, 1e-320, 5e-324, 1.7976931348623157e308]])
    rng.shuffle(vals)
    seps = [" ", "\n", "\t", ", ", " \r\n", "  "]
    fmts = ["%.17g", "%.17f", "%r", "%.3e", "%.10g"]
    toks = []
    for i, v in enumerate(vals):
        f = fmts[i % len(fmts)]
        toks.append(repr(float(v)) if f == "%r" else ((f % v) if abs(v) < 1e20 or "f" not in f else "%.17g" % v))
    text = "".join(t + seps[i % len(seps)] for i, t in enumerate(toks)).encode()
    assert len(text) > (1 << 16)  # long enough for the threaded path
    expect = np.array([float(t) for t in toks])
    for threads in (1, 3, 8):
        got, consumed, bad = _parse(H, text, len(toks) + 10, threads)
        assert bad == 0 and consumed == len(text)
        assert np.array_equal(got.view(np.uint64), expect.view(np.uint64))  # bit-exact, signed zeros and denormals included
        # a block that fills up in the middle: exactly max values, and the rest of the text parses to the remainder
        for cap in (1, 777, len(toks) // 2, len(toks) - 1, len(toks)):
            got, consumed, bad = _parse(H, text, cap, threads)
            assert len(got) == cap and bad == 0
            assert np.array_equal(got.view(np.uint64), expect[:cap].view(np.uint64))
            rest, c2, _ = _parse(H, text[consumed:], len(toks), threads)
            assert np.array_equal(rest.view(np.uint64), expect[cap:].view(np.uint64))
    # the exact fast conversion (Eisel-Lemire) behind the parser, against Python's correctly rounded float() on shapes
    # that stress it: 1-19 digit significands with exponents over the whole range, ties above 2^53, subnormals,
    # overflow, more than 19 digits, leading zeros, signs
    toks2 = []
    for i in range(120000):
        kind = i % 8
        if kind == 0:
            toks2.append("%de%d" % (rng.integers(1, 2 ** 63), rng.integers(-345, 330)))
        elif kind == 1:
            toks2.append("%d" % (2 ** 53 + 2 * int(rng.integers(0, 4096)) + 1))
        elif kind == 2:
            toks2.append("%.*e" % (int(rng.integers(0, 19)), rng.uniform(1, 10) * 10.0 ** int(rng.integers(-320, 308))))
        elif kind == 3:
            toks2.append("%s0.%s%d" % ("-" if i & 8 else "", "0" * int(rng.integers(0, 30)), rng.integers(1, 2 ** 62)))
        elif kind == 4:
            toks2.append(repr(float(np.float64(rng.integers(1, 2 ** 62)).view(np.float64)) if False else float(rng.standard_cauchy())))
        elif kind == 5:
            toks2.append("%d.%d" % (rng.integers(0, 10 ** 6), rng.integers(0, 10 ** 13)))
        elif kind == 6:
            toks2.append("%d%d" % (rng.integers(1, 2 ** 62), rng.integers(1, 2 ** 62)))  # > 19 digits: strtod path
        else:
            toks2.append("%.17g" % (float(np.frombuffer(rng.bytes(8), dtype=np.uint64)[0] % (2 ** 62)) * 4.9e-324))
    toks2 += ["1e400", "-1e400", "1e-400", "4.9e-324", "2.2250738585072014e-308", "2.2250738585072011e-308", "1.7976931348623157e308",
              "9007199254740993", "9007199254740992.5", "0e0", "-0.0", "+1.5", "1.e5", ".5e1", "00012.500", "1E+2"]
    text2 = (" ".join(toks2) + "\n").encode()
    expect2 = np.array([float(t) for t in toks2])
    for threads in (1, 8):
        got, consumed, bad = _parse(H, text2, len(toks2) + 1, threads)
        assert bad == 0 and len(got) == len(toks2)
        diff = np.nonzero(got.view(np.uint64) != expect2.view(np.uint64))[0]
        assert len(diff) == 0, [(toks2[k], got[k], expect2[k]) for k in diff[:5]]
    # a token that is not a number ends the conversion there (the reference's fscanf stops too)
    broken = b"1.5 2.5\n3.5 oops 4.5\n" + text
    for threads in (1, 8):
        got, consumed, bad = _parse(H, broken, 100000, threads)
        assert bad == 1 and list(got) == [1.5, 2.5, 3.5]
    # short texts, leading / trailing separators, empty input
    got, consumed, bad = _parse(H, b"  \n 7 ,8\n\n", 10, 8)
    assert list(got) == [7.0, 8.0] and consumed == 10 and bad == 0
    got, consumed, bad = _parse(H, b"", 10, 8)
    assert len(got) == 0 and bad == 0


def test_fast_formatter_matches_printf():
    """the output side of the stream: emub_fast_format17 gives the bytes of printf("%.17f\\n") (exact expansion, ties to
    even) or declines"""
    from madaiemulator_b200 import engine
    H = engine.host_lib()
    H.emub_fast_format17.restype = ctypes.c_int
    H.emub_fast_format17.argtypes = [ctypes.c_double, ctypes.c_char_p]
    rng = np.random.default_rng(2)
    vals = list(rng.uniform(-3, 3, 20000)) + list(rng.normal(size=5000) * 1e-12) + list(rng.normal(size=5000) * 1e12) + \
        [(2 * int(k) + 1) / 262144.0 for k in rng.integers(0, 10 ** 5, 5000)] + \
        [0.0, -0.0, 5e-324, -5e-324, 1e-18, 0.5e-17, 1.5e-17, 0.99999999999999999, 9.9999999999999999, 2.0 ** 62, -(2.0 ** 62), 123456789.125,
         0.000000000000000005, 1.0000000000000002]
    buf = ctypes.create_string_buffer(64)
    for v in vals:
        n = H.emub_fast_format17(float(v), buf)
        assert n > 0 and buf.raw[:n] == ("%.17f\n" % v).encode(), (v, buf.raw[:n])
    for v in (float("inf"), float("-inf"), float("nan"), 2.0 ** 63, -1e300):
        assert H.emub_fast_format17(v, buf) == 0


def _compare_protocol(out_text, golden_path, nt, nheader, nr=None):
    got = out_text.split("\n")
    ref = open(golden_path).read().split("\n")
    assert len(got) == len(ref)
    assert got[:nheader] == ref[:nheader]  # header lines are byte-identical
    g = np.array(got[nheader:-1], dtype=np.float64).reshape(-1, nt, 2)
    r = np.array(ref[nheader:-1], dtype=np.float64).reshape(-1, nt, 2)
    if nr is not None:  # --pca_output: only the first nr entries are meaningful (interactive_emulator.c:426-429)
        g, r = g[:, :nr], r[:, :nr]
    assert np.max(np.abs(g[..., 0] - r[..., 0]) / np.maximum(1.0, np.abs(r[..., 0]))) < 1e-9
    assert np.max(np.abs(g[..., 1] - r[..., 1])) < 1e-9 * max(1.0, np.max(np.abs(r[..., 1])))
    # same text format: 17 digits after the point
    assert all(len(x.split(".")[1]) == 17 for x in got[nheader:nheader + 20])


@pytest.mark.gpu
def test_streaming_interactive_mode_matches_reference_cli():
    assert os.path.exists(TOOL), "run make -C madaiemulator_b200/host"
    for name, pts, nt, d in (("uni-simple-o1", "uni-simple.points", 1, 1), ("multi-simple-o0", "multi-simple.points", 6, 3)):
        snap = os.path.join(CLI_DIR, name + ".snapshot")
        inp = open(os.path.join(CLI_DIR, pts), "rb").read()
        out = subprocess.run([TOOL, "interactive_mode", snap], input=inp, capture_output=True, check=True, timeout=300).stdout.decode()
        _compare_protocol(out, os.path.join(CLI_DIR, name + ".interactive.txt"), nt, 1 + d + 1 + 2 * nt)
    snap = os.path.join(CLI_DIR, "multi-simple-o0.snapshot")
    inp = open(os.path.join(CLI_DIR, "multi-simple.points"), "rb").read()
    out = subprocess.run([TOOL, "interactive_mode", snap, "--quiet"], input=inp, capture_output=True, check=True, timeout=300).stdout.decode()
    _compare_protocol(out, os.path.join(CLI_DIR, "multi-simple-o0.interactive_quiet.txt"), 6, 0)
    out = subprocess.run([TOOL, "interactive_mode", snap, "--pca_output"], input=inp, capture_output=True, check=True, timeout=300).stdout.decode()
    _compare_protocol(out, os.path.join(CLI_DIR, "multi-simple-o0.interactive_pca.txt"), 6, 0, nr=5)
    # tiny blocks give the same bytes as one big block
    small = subprocess.run([TOOL, "interactive_mode", snap, "--quiet", "--block", "7"], input=inp, capture_output=True, check=True, timeout=300).stdout.decode()
    big = subprocess.run([TOOL, "interactive_mode", snap, "--quiet"], input=inp, capture_output=True, check=True, timeout=300).stdout.decode()
    assert small == big


@pytest.mark.gpu
def test_binary_framing():
    """--binary: the framing the reference's BINARY_INTERACTIVE_MODE switch describes (interactive_emulator.c:119,
    :418-436): d raw doubles per point in, (mean_i, variance_i) raw doubles per observable out, the text header kept
    unless --quiet.  The reference's own binary build cannot make a fixture: its loop compares fread's item count (1)
    with expected_r = sizeof(double) (:391-392, :423) and leaves before the first answer -- checked here by building
    it.  So the values are compared with the reference CLI's TEXT output for the same points (1e-9), bit for bit
    with the tool's own text path up to the 17 decimals that path prints, for ragged reads and every block size."""
    snap = os.path.join(CLI_DIR, "multi-simple-o0.snapshot")
    nt, d = 6, 3
    pts = np.array(open(os.path.join(CLI_DIR, "multi-simple.points")).read().split(), dtype=np.float64)
    raw = pts.tobytes()
    out = subprocess.run([TOOL, "interactive_mode", snap, "--quiet", "--binary"], input=raw, capture_output=True, check=True, timeout=300).stdout
    assert len(out) == (len(pts) // d) * 2 * nt * 8
    got = np.frombuffer(out, dtype=np.float64).reshape(-1, nt, 2)
    ref = np.array(open(os.path.join(CLI_DIR, "multi-simple-o0.interactive_quiet.txt")).read().split(), dtype=np.float64).reshape(-1, nt, 2)
    assert np.max(np.abs(got[..., 0] - ref[..., 0]) / np.maximum(1.0, np.abs(ref[..., 0]))) < 1e-9
    assert np.max(np.abs(got[..., 1] - ref[..., 1])) < 1e-9 * max(1.0, np.max(np.abs(ref[..., 1])))
    text = subprocess.run([TOOL, "interactive_mode", snap, "--quiet"], input=open(os.path.join(CLI_DIR, "multi-simple.points"), "rb").read(),
                          capture_output=True, check=True, timeout=300).stdout.decode()
    assert text == "".join("%.17f\n" % v for v in got.ravel())
    # block size does not change a bit; a trailing partial point (fewer than d doubles) is dropped like the reference's
    # short fread (:423-424)
    small = subprocess.run([TOOL, "interactive_mode", snap, "--quiet", "--binary", "--block", "7"], input=raw + b"\x00" * 12,
                           capture_output=True, check=True, timeout=300).stdout
    assert small == out
    # with the header: text lines first (interactive_emulator.c:398-414), then raw doubles
    full = subprocess.run([TOOL, "interactive_mode", snap, "--binary"], input=raw[:3 * d * 8], capture_output=True, check=True, timeout=300).stdout
    header = "".join(["%d\n" % d] + ["param_%d\n" % i for i in range(d)] + ["%d\n" % (2 * nt)] +
                     ["mean_%d\nvariance_%d\n" % (i, i) for i in range(nt)]).encode()
    assert full[:len(header)] == header
    assert full[len(header):] == out[:3 * 2 * nt * 8]
    # pca_output: first nr entries of each row as in the text protocol
    pca = subprocess.run([TOOL, "interactive_mode", snap, "--pca_output", "--binary"], input=raw, capture_output=True, check=True, timeout=300).stdout
    gp = np.frombuffer(pca, dtype=np.float64).reshape(-1, nt, 2)[:, :5]
    rp = np.array(open(os.path.join(CLI_DIR, "multi-simple-o0.interactive_pca.txt")).read().split(), dtype=np.float64).reshape(-1, nt, 2)[:, :5]
    assert np.max(np.abs(gp - rp)) < 1e-9 * max(1.0, np.max(np.abs(rp)))
    # empty input
    assert subprocess.run([TOOL, "interactive_mode", snap, "--quiet", "--binary"], input=b"", capture_output=True, check=True, timeout=300).stdout == b""


@pytest.mark.gpu
def test_request_response_client_is_served_point_by_point():
    """A client that writes one point and waits for its 2*nt answer lines (the reference flushes per point) must not
    dead-lock on the block reader."""
    snap = os.path.join(CLI_DIR, "multi-simple-o0.snapshot")
    p = subprocess.Popen([TOOL, "interactive_mode", snap, "--quiet"], stdin=subprocess.PIPE, stdout=subprocess.PIPE)
    ref = open(os.path.join(CLI_DIR, "multi-simple-o0.interactive_quiet.txt")).read().split("\n")
    pts = open(os.path.join(CLI_DIR, "multi-simple.points")).read().split("\n")
    for q in range(3):
        p.stdin.write((pts[q] + "\n").encode())
        p.stdin.flush()
        lines = [p.stdout.readline().decode().strip() for _ in range(12)]
        vals = np.array(lines, dtype=np.float64)
        assert np.max(np.abs(vals - np.array(ref[12 * q:12 * q + 12], dtype=np.float64))) < 1e-9
    p.stdin.close()
    assert p.wait(timeout=60) == 0


@pytest.mark.gpu
def test_reference_cli_on_the_engine():
    """The reference's own interactive_emulator.c, unmodified, linked with integration/libemu_glue.c."""
    if not os.path.exists(DROPIN_CLI):
        pytest.skip("oracle/_ref/interactive_emulator_dropin not built")
    snap = os.path.join(CLI_DIR, "multi-simple-o0.snapshot")
    inp = open(os.path.join(CLI_DIR, "multi-simple.points"), "rb").read()
    out = subprocess.run([DROPIN_CLI, "interactive_mode", snap], input=inp, capture_output=True, check=True, timeout=300).stdout.decode()
    _compare_protocol(out, os.path.join(CLI_DIR, "multi-simple-o0.interactive.txt"), 6, 1 + 3 + 1 + 12)


@pytest.mark.gpu
def test_query_blocks_shard_over_two_gpus():
    """Query points are independent: with --devices 0,1 every block is split between two device
    replicas (one host thread per GPU) and the output is byte-identical to the single-GPU run."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    snap = os.path.join(CLI_DIR, "multi-simple-o0.snapshot")
    inp = open(os.path.join(CLI_DIR, "multi-simple.points"), "rb").read()
    one = subprocess.run([TOOL, "interactive_mode", snap, "--quiet"], input=inp, capture_output=True, check=True, timeout=300).stdout
    two = subprocess.run([TOOL, "interactive_mode", snap, "--quiet", "--devices", "0,1"], input=inp, capture_output=True, check=True, timeout=300).stdout
    assert one == two
