"""Shared helpers for the parity tests: golden-fixture loading and error measures."""
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def unhex(v):
    if isinstance(v, str):
        return float.fromhex(v)
    return np.array([unhex(x) for x in v], dtype=np.float64)


_FLOAT_KEYS = ("X", "y", "theta_full", "cov_row_values", "cov_diag", "cov_sum", "H", "ranges", "pts", "emu_mean",
               "emu_var", "emu_beta", "kappa", "theta_less_amp", "negL_logsum", "logdet", "sigma2", "beta",
               "negL_literal", "grad", "deriv2_row0", "cinv_trace", "cinv_row0")


def load_golden(name):
    with open(os.path.join(GOLDEN_DIR, name + ".json")) as f:
        raw = json.load(f)
    c = dict(raw)
    for k in _FLOAT_KEYS:
        if k in raw:
            c[k] = unhex(raw[k])
    c["X"] = c["X"].reshape(c["n"], c["d"])
    c["pts"] = c["pts"].reshape(-1, c["d"])
    return c


def golden_names():
    with open(os.path.join(GOLDEN_DIR, "index.json")) as f:
        return json.load(f)["cases"]


def load_all_golden():
    return {n: load_golden(n) for n in golden_names()}


def relerr(a, b, floor=0.0):
    """max |a-b| / max(|b|, floor)"""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(b), floor)
    den = np.where(den == 0, 1.0, den)
    return float(np.max(np.abs(a - b) / den))


def scaled_err(a, b, scale):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / scale)
