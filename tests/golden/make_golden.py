#!/usr/bin/env python
"""Generate tests/golden/*.json from the REFERENCE's own sources.

Runs only in the build container (needs /root/reference): the reference's hot-path C files are
compiled unmodified against oracle/gsl_shim (oracle/Makefile -> oracle/_ref/libemu_ref.so) and
driven through oracle/ref_driver.c.  Inputs come from the reference's shipped example data
(test/uni-simple, test/uni-2d-param, test/multi-simple) plus one synthetic d=10 design; outputs are
what the reference functions return at fixed hyper-parameters.  All doubles are stored as C99 hex
floats so the fixtures are bit-exact.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from madaiemulator_b200 import datasets as ds  # noqa: E402
from oracle.pyoracle import RefOracle, build  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def hx(a):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 0:
        return float(a).hex()
    return [hx(x) for x in a]


def read_points(path, d, limit):
    v = np.array(open(path).read().split(), dtype=np.float64)
    pts = v[: (len(v) // d) * d].reshape(-1, d)
    return pts[:limit].copy()


def case(name, X, y, kernel, order, theta_full, pts):
    r = RefOracle(X, y, kernel, order)
    n, d = X.shape
    out = dict(name=name, kernel=kernel, order=order, n=n, d=d, X=hx(X), y=hx(y), theta_full=hx(theta_full))
    C = r.cov_matrix(theta_full)
    rows = sorted(set([0, n // 3, n // 2, n - 1]))
    out["cov_rows"] = rows
    out["cov_row_values"] = hx(C[rows])
    out["cov_diag"] = hx(np.diag(C))
    out["cov_sum"] = hx(C.sum())
    out["H"] = hx(r.h_matrix())
    out["ranges"] = hx(r.ranges())
    out["pts"] = hx(pts)
    e = r.emulator(theta_full)
    mean, var = e.emulate(pts)
    out["emu_mean"] = hx(mean)
    out["emu_var"] = hx(var)
    out["emu_beta"] = hx(e.beta())
    out["kappa"] = hx(r.cov_pair(pts[0], pts[0], theta_full))
    if kernel == 1:
        th = np.asarray(theta_full[1:], dtype=np.float64)
        out["theta_less_amp"] = hx(th)
        ls = r.eval_logsum(th)
        out["negL_logsum"] = hx(ls["negL"])
        out["logdet"] = hx(ls["logdet"])
        out["sigma2"] = hx(ls["sigma2"])
        out["beta"] = hx(ls["beta"])
        out["negL_literal"] = hx(r.eval(th))  # evalFnMulti as shipped (determinant running product)
        out["grad"] = hx(r.grad(th))          # gradFnMulti
        D = r.deriv_matrix(th[1], 2)          # derivative_l_gauss for the first length
        out["deriv2_row0"] = hx(D[0])
        rc, Cinv = r.cinverse(th)
        out["cinv_trace"] = hx(np.trace(Cinv))
        out["cinv_row0"] = hx(Cinv[0])
    return out


def main():
    build(ref=True)
    os.makedirs(OUT, exist_ok=True)
    fixtures = []
    Xu, Yu = ds.load_input_model_file(f"{REF}/test/uni-simple/input_model_file.dat")
    pu = read_points(f"{REF}/test/uni-simple/sample_locations.dat", 1, 100)
    X2, Y2 = ds.load_input_model_file(f"{REF}/test/uni-2d-param/Latin_square_sampling_2d_samp_fn_200.dat")
    p2 = read_points(f"{REF}/test/uni-2d-param/sample_locations.dat", 2, 64)
    Xm, Ym = ds.load_input_model_file(f"{REF}/test/multi-simple/multi-test-input.dat")
    pm = np.array([[0.5, 1.0, 1.2], [0.0, 0.0, 0.0], Xm[7], Xm[7] + 1e-12, Xm[50] * 0.5])

    cases = []
    for order in (0, 1, 2, 3):
        cases.append(case(f"uni-simple-o{order}", Xu, Yu[:, 0].copy(), 1, order, [-1.8, -3.0, -0.5], pu))
    for order in (0, 2):
        cases.append(case(f"uni-2d-o{order}", X2, Y2[:, 0].copy(), 1, order, [-2.0, -4.0, -1.0, -0.8], p2))
    pca = ds.pca_decompose(Ym, 0.99)
    for comp in range(min(2, pca["nr"])):
        for order in (0, 1):
            cases.append(case(f"multi-simple-pc{comp}-o{order}", Xm, pca["Z"][:, comp].copy(), 1, order,
                              [0.1, -3.5, 0.2, 0.4, 0.0], pm))
    # Matern kernels at function level + prediction (raw amplitude / raw nugget, Q1)
    for kern, nm in ((2, "m32"), (3, "m52")):
        cases.append(case(f"uni-simple-{nm}", Xu, Yu[:, 0].copy(), kern, 1, [0.8, 0.01, -0.3], pu[:40]))
        cases.append(case(f"multi-simple-{nm}", Xm, pca["Z"][:, 0].copy(), kern, 0, [1.3, 0.02, 0.5], pm))
    # synthetic d = 10 (bench design family), n = 256
    Xs = ds.synthetic_design(256, 10)
    ys = ds.synthetic_response(Xs)
    ps = ds.synthetic_queries(32, 10)
    ps[1] = Xs[17]
    ths = np.concatenate([[-0.7, -4.0], np.full(10, 1.0)])
    cases.append(case("synthetic-n256-d10-o1", Xs, ys, 1, 1, ths, ps))
    for c in cases:
        with open(os.path.join(OUT, c["name"] + ".json"), "w") as f:
            json.dump(c, f)
        fixtures.append(c["name"])
    # PCA / back-projection fixture for multi-simple (gen_pca_decomp + emulate_point_multi)
    with open(os.path.join(OUT, "index.json"), "w") as f:
        json.dump(dict(cases=fixtures, generator="tests/golden/make_golden.py",
                       source="reference sources compiled unmodified against oracle/gsl_shim"), f, indent=1)
    print("wrote", len(fixtures), "fixtures")


if __name__ == "__main__":
    main()
