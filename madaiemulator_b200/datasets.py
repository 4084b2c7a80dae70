"""Inputs for the hot path: the reference's INPUT_MODEL_FILE reader and the synthetic designs
that bench.py / the parity tests use.

Nothing here touches the GPU; it is host-side data preparation only.
"""
import numpy as np

SEED = 20261018
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(z):
    """Counter-based generator: splitmix64 finaliser applied to uint64 counters (vectorised)."""
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        z = z ^ (z >> np.uint64(31))
    return z


def uniform01(seed, counters):
    """53-bit uniform in [0,1) from splitmix64(seed * 2^40 + counter)."""
    with np.errstate(over="ignore"):
        base = (np.uint64(seed) << np.uint64(40)) & _M64
        z = splitmix64(base + np.asarray(counters, dtype=np.uint64))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def synthetic_design(n, d, seed=SEED, lo=-2.5, hi=2.5, row0=0):
    """n x d design, x_ik = lo + (hi-lo) u(seed, i*d + k); box [-2.5, 2.5]^d as the reference's
    sensitivity-analysis sampler (src/libSA/make-samples.R:21)."""
    idx = (np.arange(row0, row0 + n, dtype=np.uint64)[:, None] * np.uint64(d)
           + np.arange(d, dtype=np.uint64)[None, :])
    return lo + (hi - lo) * uniform01(seed, idx)


def _coeffs(d, seed, rot):
    c = uniform01(seed + 7919, np.arange(3 * 15 + 15 * 15, dtype=np.uint64) + np.uint64(1000 * rot))
    a1 = (2 * c[0:15] - 1)[:d]
    a2 = (2 * c[15:30] - 1)[:d]
    a3 = (2 * c[30:45] - 1)[:d]
    M = (c[45:].reshape(15, 15) - 0.5)[:d, :d] * (2.0 / d)
    return a1, a2, a3, M


def synthetic_response(X, output=0, seed=SEED, standardise=True):
    """Oakley-O'Hagan-form response y = a1.x + a2.sin x + a3.cos x + x^T M x
    (functional form of src/libSA/make-samples.R:24-26; the coefficients are drawn from the
    counter generator, rotated per output), standardised to zero mean / unit variance."""
    X = np.asarray(X, dtype=np.float64)
    a1, a2, a3, M = _coeffs(X.shape[1], seed, output)
    y = X @ a1 + np.sin(X) @ a2 + np.cos(X) @ a3 + np.einsum("ij,jk,ik->i", X, M, X)
    if standardise:
        y = (y - y.mean()) / y.std()
    return y


def synthetic_model(n, d, nt=1, seed=SEED):
    X = synthetic_design(n, d, seed)
    Y = np.stack([synthetic_response(X, t, seed) for t in range(nt)], axis=1)
    return X, Y


def synthetic_queries(m, d, seed=SEED + 1, row0=0):
    return synthetic_design(m, d, seed, row0=row0)


def default_theta_less_amp(d, kernel=1):
    """Fixed evaluation point of SURVEY 8d: log-nugget -4, log-lengths 1.0."""
    if kernel in (2, 3):
        return np.array([-4.0, 1.0])
    return np.concatenate([[-4.0], np.full(d, 1.0)])


def load_input_model_file(path):
    """INPUT_MODEL_FILE: nt, d, n, then X (n*d) and Y (n*nt), whitespace separated
    (reference src/interactive_emulator.c:212-245)."""
    tok = open(path).read().split()
    nt, d, n = int(tok[0]), int(tok[1]), int(tok[2])
    v = np.array(tok[3:3 + n * d + n * nt], dtype=np.float64)
    X = v[:n * d].reshape(n, d).copy()
    Y = v[n * d:].reshape(n, nt).copy()
    return X, Y


def pca_decompose(Y, vfrac=0.99):
    """PCA of the training outputs as gen_pca_decomp does (src/multi_modelstruct.c:172-338):
    centre by column means, Sigma = Yc^T Yc / n, eigen-decomposition sorted descending, keep
    nr = min(k*+1, nt-1) components (loop :267-272), Z = Yc U_r diag(lambda^-1/2).
    Returns dict(mean, evals, evecs (nt x nr), Z (n x nr), nr)."""
    Y = np.asarray(Y, dtype=np.float64)
    n, nt = Y.shape
    mean = Y.mean(axis=0)
    Yc = Y - mean
    cov = (Yc.T @ Yc) / n
    w, V = np.linalg.eigh(cov)
    order = np.argsort(-w)
    w, V = w[order], V[:, order]
    total = w.sum()
    frac, i = 0.0, 0
    while frac < vfrac and (i + 1) < nt:
        frac = w[:i].sum() / total
        i += 1
    nr = i
    if nt == 1:
        nr = 1
    evals, evecs = w[:nr].copy(), V[:, :nr].copy()
    Z = (Yc @ evecs) / np.sqrt(evals)[None, :]
    return dict(mean=mean, evals=evals, evecs=evecs, Z=Z, nr=nr)
