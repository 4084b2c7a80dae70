"""n=4096, d=10 scalar-GP snapshot for tools/cfg5_stream.sh (synthetic design of SURVEY 8d, fixed thetas)."""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madaiemulator_b200 import datasets as ds  # noqa: E402
from madaiemulator_b200 import engine  # noqa: E402

n, d = 4096, 10
X = np.ascontiguousarray(ds.synthetic_design(n, d))
y = np.ascontiguousarray(ds.synthetic_response(X)).reshape(n, 1)
th = np.ascontiguousarray(np.concatenate([[0.0, -4.0], np.full(d, 1.0)]).reshape(1, d + 2))
H = engine.host_lib()
_dp = ctypes.POINTER(ctypes.c_double)
H.emub_snapshot_from_arrays.restype = ctypes.c_void_p
H.emub_snapshot_from_arrays.argtypes = [_dp, ctypes.c_int, ctypes.c_int, _dp, ctypes.c_int, _dp, ctypes.c_int, ctypes.c_int, ctypes.c_int]
H.emub_snapshot_save_path.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
H.emub_snapshot_free.argtypes = [ctypes.c_void_p]
sp = H.emub_snapshot_from_arrays(X.ctypes.data_as(_dp), n, d, y.ctypes.data_as(_dp), 1, th.ctypes.data_as(_dp), d + 2, 1, 0)
assert sp
assert H.emub_snapshot_save_path(sp, sys.argv[1].encode()) == 0
H.emub_snapshot_free(sp)
print("wrote", sys.argv[1])
