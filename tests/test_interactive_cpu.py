"""CPU: the streaming interactive_mode layer (madaiemulator_b200/host/emub_interactive.c + emub_cli.c + the snapshot
loader and the number formats) linked against a mock of the C-ABI entry points it calls (tests/mock/mock_predict.c: an
analytic "emulator" the test recomputes in the same operation order), so that the protocol of
interactive_emulator.c:369-450 -- text and binary framing, header lines, block boundaries, bad input, the
request / response pattern, the split over devices -- is checked byte for byte without a GPU."""
import math
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "madaiemulator_b200", "host")
CLI_DIR = os.path.join(ROOT, "tests", "golden", "cli")
SNAP = os.path.join(CLI_DIR, "multi-simple-o0.snapshot")
NT, NR, D, N = 6, 5, 3, 100


@pytest.fixture(scope="module")
def cli(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("mockcli") / "cli")
    subprocess.check_call(["gcc", "-std=gnu99", "-O1", "-ffp-contract=off", "-o", out] +
                          [os.path.join(HOST, f) for f in ("emub_cli.c", "emub_interactive.c", "emub_snapshot.c", "emub_fastfloat.c")] +
                          [os.path.join(ROOT, "tests", "mock", "mock_predict.c"), "-I" + os.path.join(ROOT, "include"), "-lm", "-lpthread"])
    return out


def _theta1_per_component():
    import ctypes
    from tests.test_interactive_stream import _Snap
    from madaiemulator_b200 import engine
    H = engine.host_lib()
    H.emub_snapshot_load_path.restype = ctypes.POINTER(_Snap)
    H.emub_snapshot_load_path.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
    H.emub_snapshot_free.argtypes = [ctypes.POINTER(_Snap)]
    H.emub_snapshot_free.restype = None
    err = ctypes.create_string_buffer(256)
    sp = H.emub_snapshot_load_path(SNAP.encode(), err, 256)
    s = sp.contents
    th1 = [s.components[c].thetas[1] for c in range(NR)]
    mean = [s.training_mean[i] for i in range(NT)]
    evals = [s.pca_evals_r[j] for j in range(NR)]
    evecs = [[s.pca_evecs_r[i * NR + j] for j in range(NR)] for i in range(NT)]
    H.emub_snapshot_free(sp)
    return th1, mean, evals, evecs


def _expected(points, pca):
    """(mean, var) rows the mock produces, same operations in the same order, Python floats (IEEE double)"""
    th1, tmean, evals, evecs = _theta1_per_component()
    rows = []
    for x in points:
        s = 0.0
        for k in range(D):
            s += x[k] * float(k + 1)
        pm = [s * float(j + 1) + th1[j] for j in range(NR)]
        pv = [0.5 * float(j + 1) + x[0] * x[0] for j in range(NR)]
        if pca:
            rows.append((pm + [0.0] * (NT - NR), pv + [0.0] * (NT - NR)))
            continue
        m, v = [], []
        for i in range(NT):
            a, b = tmean[i], 0.0
            for j in range(NR):
                a += evecs[i][j] * math.sqrt(evals[j]) * pm[j]
                b += evecs[i][j] * evecs[i][j] * evals[j] * pv[j]
            m.append(a)
            v.append(b)
        rows.append((m, v))
    return rows


def _text(rows, count):
    out = []
    for m, v in rows:
        for i in range(count):
            out.append("%.17f\n" % m[i])
            out.append("%.17f\n" % v[i])
    return "".join(out)


def _points():
    return np.array(open(os.path.join(CLI_DIR, "multi-simple.points")).read().split(), dtype=np.float64).reshape(-1, D)


def _run(cli, args, data, **kw):
    return subprocess.run([cli, "interactive_mode", SNAP] + args, input=data, capture_output=True, timeout=120, **kw)


def test_text_protocol_byte_for_byte(cli):
    pts = _points()
    inp = open(os.path.join(CLI_DIR, "multi-simple.points"), "rb").read()
    want = _text(_expected(pts.tolist(), False), NT)
    r = _run(cli, ["--quiet"], inp)
    assert r.returncode == 0 and r.stdout.decode() == want
    # the header of the non-quiet mode is the reference's (interactive_emulator.c:398-414), then the same answers
    nheader = 1 + D + 1 + 2 * NT
    golden = open(os.path.join(CLI_DIR, "multi-simple-o0.interactive.txt")).read().split("\n")
    r = _run(cli, [], inp)
    got = r.stdout.decode()
    assert got.split("\n")[:nheader] == golden[:nheader]
    assert "\n".join(got.split("\n")[nheader:]) == want
    # --pca_output: nr meaningful (mean, variance) pairs per point in PCA space, quiet implied (the reference's fall-through, :589-601)
    r = _run(cli, ["--pca_output"], inp)
    rows = _expected(pts.tolist(), True)
    got = r.stdout.decode().split("\n")
    per_point = len(got[:-1]) // len(pts)
    assert per_point * len(pts) == len(got) - 1
    for q in range(len(pts)):
        for j in range(NR):
            assert got[q * per_point + 2 * j] == "%.17f" % rows[q][0][j] and got[q * per_point + 2 * j + 1] == "%.17f" % rows[q][1][j]


@pytest.mark.parametrize("block", [1, 7, 64, 100000])
def test_block_size_does_not_change_a_byte(cli, block):
    inp = open(os.path.join(CLI_DIR, "multi-simple.points"), "rb").read()
    want = _text(_expected(_points().tolist(), False), NT)
    assert _run(cli, ["--quiet", "--block", str(block)], inp).stdout.decode() == want


def test_separators_partial_points_and_bad_tokens(cli):
    pts = [[0.25, 0.5, 0.75], [1e-3, 2.5e1, -3.0], [0.1, 0.2, 0.3]]
    want3 = _text(_expected(pts, False), NT)
    want2 = _text(_expected(pts[:2], False), NT)
    # fscanf("%lf%*c") accepts any single separator character after a number (:418-423): blanks, tabs, CR LF, commas
    r = _run(cli, ["--quiet"], b"0.25,0.5\t0.75\r\n1e-3 2.5e1\n-3.0\n0.1 0.2 0.3")
    assert r.stdout.decode() == want3
    # an incomplete last point is not answered
    r = _run(cli, ["--quiet"], b"0.25 0.5 0.75\n1e-3 25 -3\n0.1 0.2")
    assert r.stdout.decode() == want2
    # a token that is not a number ends the session after the points read so far (the reference's loop leaves at the
    # first failed conversion, :419-422)
    for bad in (b"0.25 0.5 0.75\n1e-3 25 -3\nquit\n0.1 0.2 0.3\n", b"0.25 0.5 0.75\n1e-3 25 -3\n0.1 0.2abc 0.3\n"):
        r = _run(cli, ["--quiet"], bad)
        assert r.stdout.decode() == want2
    # nothing in, nothing out
    assert _run(cli, ["--quiet"], b"").stdout == b""


def test_binary_framing(cli):
    """BINARY_INTERACTIVE_MODE (interactive_emulator.c:119-135, :424-441): raw doubles in, raw doubles out"""
    pts = _points()
    raw = pts.tobytes()
    rows = _expected(pts.tolist(), False)
    want = b"".join(struct.pack("<2d", m[i], v[i]) for m, v in rows for i in range(NT))
    r = _run(cli, ["--quiet", "--binary"], raw)
    assert r.returncode == 0 and r.stdout == want
    # trailing bytes that are not a whole point are ignored; tiny blocks give the same bytes
    r = _run(cli, ["--quiet", "--binary", "--block", "7"], raw + b"\x00" * 12)
    assert r.stdout == want
    assert _run(cli, ["--quiet", "--binary"], b"").stdout == b""


def test_devices_split_blocks_without_changing_the_stream(cli):
    inp = open(os.path.join(CLI_DIR, "multi-simple.points"), "rb").read() * 40
    one = _run(cli, ["--quiet", "--block", "64"], inp).stdout
    for devs in ("0,1", "0,1,2", "3,2,1,0,4,5,6,7"):
        assert _run(cli, ["--quiet", "--block", "64", "--devices", devs], inp).stdout == one
    assert len(one.split(b"\n")) - 1 == 2 * NT * (len(inp.split()) // D)


def test_request_response_client_gets_each_answer_before_the_next_question(cli):
    """the per-point pattern of an MCMC driver on a pipe: write one point, read its 2 nt lines, only then write the next"""
    pts = [[0.1 * i, 0.2, 0.3 + 0.01 * i] for i in range(1, 6)]
    rows = _expected(pts, False)
    p = subprocess.Popen([cli, "interactive_mode", SNAP, "--quiet"], stdin=subprocess.PIPE, stdout=subprocess.PIPE)
    try:
        for x, (m, v) in zip(pts, rows):
            p.stdin.write((" ".join(repr(c) for c in x) + "\n").encode())
            p.stdin.flush()
            got = [p.stdout.readline().decode() for _ in range(2 * NT)]
            want = [("%.17f\n" % (m[i // 2] if i % 2 == 0 else v[i // 2])) for i in range(2 * NT)]
            assert got == want
        p.stdin.close()
        assert p.wait(timeout=30) == 0
    finally:
        if p.poll() is None:
            p.kill()
