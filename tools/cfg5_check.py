"""checks the streamed answers of tools/cfg5_stream.sh: line counts, text == binary, and a sample against the C-ABI."""
import os
import subprocess
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madaiemulator_b200 import engine  # noqa: E402

W, npts = sys.argv[1], int(sys.argv[2])
out_bin = np.fromfile(os.path.join(W, "out.bin")).reshape(-1, 2)
assert out_bin.shape[0] == npts, out_bin.shape
nlines = int(subprocess.check_output(["wc", "-l", os.path.join(W, "out.txt")]).split()[0])
assert nlines == 2 * npts, nlines
K = 20000
head = np.array(subprocess.check_output(["head", "-n", str(2 * K), os.path.join(W, "out.txt")]).split(), dtype=np.float64).reshape(-1, 2)
tail = np.array(subprocess.check_output(["tail", "-n", str(2 * K), os.path.join(W, "out.txt")]).split(), dtype=np.float64).reshape(-1, 2)
# "%.17f" keeps 17 decimals: text agrees with the binary answers to that
print("text vs binary, head/tail max abs diff:", np.max(np.abs(head - out_bin[:K])), np.max(np.abs(tail - out_bin[-K:])))
assert np.max(np.abs(head - out_bin[:K])) < 1e-16 * 10 and np.max(np.abs(tail - out_bin[-K:])) < 1e-16 * 10
# the single-IO-thread run gives the same bytes as the threaded one
one = subprocess.check_output(["head", "-n", str(2 * K), os.path.join(W, "out1.txt")])
assert one == subprocess.check_output(["head", "-n", str(2 * K), os.path.join(W, "out.txt")])
# sample against the C-ABI on the snapshot's own contents
tok = open(os.path.join(W, "cfg5.snapshot")).read().split()
n, d = 4096, 10
X = np.array(tok[6:6 + n * d], dtype=np.float64).reshape(n, d)
th = np.array(tok[-d - (d + 2):-d], dtype=np.float64)
z = np.array(tok[-d - (d + 2) - n:-d - (d + 2)], dtype=np.float64)
ybar = np.array(tok[6 + n * d:6 + n * d + n], dtype=np.float64).sum() / n
pts = np.fromfile(os.path.join(W, "pts.bin"), count=K * d).reshape(K, d)
ctx = engine.Context(0)
m = engine.Model(ctx, X, z, engine.POWEREXP, 0, max_slots=1)
e = m.emulator(th)
mean, var = e.emulate(pts)
print("stream vs C-ABI: mean", np.max(np.abs(out_bin[:K, 0] - (ybar + mean))), "var", np.max(np.abs(out_bin[:K, 1] - var)))
assert np.max(np.abs(out_bin[:K, 0] - (ybar + mean))) < 1e-12 and np.max(np.abs(out_bin[:K, 1] - var)) < 1e-12
print("cfg5 stream check ok")
