"""Synthetic problem generators shared by the host side, the tests and bench.py.

All randomness comes from one counter-based generator (SplitMix64 addressed by
``(seed, stream, index)``), so a row shard of a matrix can be generated in place
on any rank -- or on the device by ``csrc/generate.cu``, which implements the
same bit-exact function -- without materialising the whole matrix anywhere.

The planted lasso follows the construction of the reference's
experiments/lasso/runme.jl:40-77 step by step (same formulas, same roles for
``pfactor``, ``lam``, ``rho``); only the random stream differs (Julia's own
stream depends on the Julia version).
"""
from __future__ import annotations

import numpy as np

_GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)

# stream ids (shared with csrc/generate.cu)
STREAM_YSTAR = 0
STREAM_MATRIX = 1
STREAM_ALPHA = 2
STREAM_XSTAR = 3
STREAM_AUX = 4


def _mix(z):
    z = np.asarray(z, dtype=np.uint64)
    z = (z ^ (z >> np.uint64(30))) * _M1
    z = (z ^ (z >> np.uint64(27))) * _M2
    return z ^ (z >> np.uint64(31))


def stream_key(seed: int, stream: int) -> np.uint64:
    """Key of a stream: mix(mix(seed) ^ (stream+1)*GOLD)."""
    with np.errstate(over="ignore"):
        s = _mix(np.array([seed], dtype=np.uint64))
        k = _mix(s ^ (np.array([stream + 1], dtype=np.uint64) * _GOLD))
    return k[0]


def bits64(seed: int, stream: int, index) -> np.ndarray:
    """SplitMix64 output number ``index`` of the stream (random access)."""
    idx = np.asarray(index, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = stream_key(seed, stream) + (idx + np.uint64(1)) * _GOLD
        return _mix(z)


def uniform01(seed: int, stream: int, index) -> np.ndarray:
    """U[0,1) with 53 random bits."""
    return (bits64(seed, stream, index) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def normal01(seed: int, stream: int, index) -> np.ndarray:
    """Standard normal by Box-Muller on the pair (2*index, 2*index+1)."""
    idx = np.asarray(index, dtype=np.uint64)
    u1 = uniform01(seed, stream, idx * np.uint64(2))
    u2 = uniform01(seed, stream, idx * np.uint64(2) + np.uint64(1))
    return np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2.0 * np.pi * u2)


def matrix_uniform_pm1(seed: int, m: int, n: int, row0: int = 0, rows: int | None = None) -> np.ndarray:
    """Rows [row0, row0+rows) of the m x n matrix ``rand(m, n) .* 2 .- 1``
    (lasso/runme.jl:50); element (i, j) uses index i*n + j."""
    rows = m - row0 if rows is None else rows
    i = np.arange(row0, row0 + rows, dtype=np.uint64)[:, None]
    j = np.arange(n, dtype=np.uint64)[None, :]
    return uniform01(seed, STREAM_MATRIX, i * np.uint64(n) + j) * 2.0 - 1.0


def spectral_norm_sq(A: np.ndarray, iters: int = 200, tol: float = 1e-13) -> float:
    """``opnorm(A)^2`` by power iteration on A'A from the all-ones vector."""
    v = np.ones(A.shape[1]) / np.sqrt(A.shape[1])
    lam = 0.0
    for _ in range(iters):
        w = A.T @ (A @ v)
        lam_new = float(np.linalg.norm(w))
        v = w / lam_new
        if abs(lam_new - lam) <= tol * lam_new:
            lam = lam_new
            break
        lam = lam_new
    return lam


def planted_lasso(m=400, n=1000, pfactor=5, seed=0, lam=1.0, rho=1.0):
    """The random lasso instance with a known minimiser, lasso/runme.jl:40-77.

    Returns a dict with A (m x n, Fortran order like Julia), b, x_star, y_star,
    lam, optimum.  ``Lf`` is left to the caller (``opnorm(A)^2``, :81).
    """
    p = n / pfactor                                               # :45  (a Float in Julia)
    i = np.arange(m, dtype=np.uint64)
    y_star = uniform01(seed, STREAM_YSTAR, i)                     # :48
    y_star = y_star / np.sqrt(np.dot(y_star, y_star))             # :49
    C = matrix_uniform_pm1(seed, m, n)                            # :50

    CTy = np.abs(C.T @ y_star)                                    # :52
    perm = np.argsort(-CTy, kind="stable")                        # :53 sortperm(rev=true)

    col = np.arange(n, dtype=np.uint64)
    u_alpha = uniform01(seed, STREAM_ALPHA, col)
    u_x = uniform01(seed, STREAM_XSTAR, col)

    alpha = np.zeros(n)
    for k in range(n):                                            # :56-68 (k+1 is Julia's i)
        jcol = perm[k]
        if k + 1 <= p:
            alpha[jcol] = lam / CTy[jcol]
        else:
            temp = CTy[jcol]
            if temp < 0.1 * lam:
                alpha[jcol] = lam
            else:
                alpha[jcol] = lam * u_alpha[jcol] / temp
    A = np.asfortranarray(C * alpha[None, :])                     # :69
    x_star = np.zeros(n)
    for k in range(n):                                            # :71-75
        jcol = perm[k]
        if k + 1 <= p:
            x_star[jcol] = u_x[jcol] * rho / np.sqrt(p) * np.sign(np.dot(A[:, jcol], y_star))
    b = A @ x_star + y_star                                       # :76
    optimum = np.sqrt(np.dot(y_star, y_star)) / 2 + lam * np.sum(np.abs(x_star))   # :77
    return dict(A=A, b=b, x_star=x_star, y_star=y_star, lam=float(lam), optimum=float(optimum),
                alpha=alpha, perm=perm, CTy=CTy)


def sparse_logreg(m=20242, n=47236, seed=0, nnz_lo=40, nnz_hi=112, w_density=0.01):
    """rcv1-shaped CSR design matrix with labels in {0,1} from a planted sparse
    model (config C2; the reference reads LIBSVM files, sparse_logreg/runme.jl:52).

    Row i gets k_i in [nnz_lo, nnz_hi) distinct uniform columns (mean ~ 75.5,
    0.16 % density), values U(0,1) normalised to unit row 2-norm.
    Returns (rowptr int64, colind int32, vals float64, y) with sorted columns.
    """
    rows = np.arange(m, dtype=np.uint64)
    k = nnz_lo + (uniform01(seed, STREAM_AUX, rows) * (nnz_hi - nnz_lo)).astype(np.int64)
    kmax = int(nnz_hi)
    slot = np.arange(kmax, dtype=np.uint64)[None, :]
    cols = (bits64(seed, STREAM_MATRIX, rows[:, None] * np.uint64(2 * kmax) + slot) % np.uint64(n)).astype(np.int64)
    vals = uniform01(seed, STREAM_MATRIX, rows[:, None] * np.uint64(2 * kmax) + np.uint64(kmax) + slot)
    rowptr = [0]
    ci, vv = [], []
    for r in range(m):
        c, first = np.unique(cols[r, : k[r]], return_index=True)
        v = vals[r, first] + 1e-3
        v = v / np.sqrt(np.dot(v, v))
        ci.append(c)
        vv.append(v)
        rowptr.append(rowptr[-1] + len(c))
    colind = np.concatenate(ci).astype(np.int32)
    values = np.concatenate(vv)
    rowptr = np.asarray(rowptr, dtype=np.int64)
    # planted model
    cj = np.arange(n, dtype=np.uint64)
    mask = uniform01(seed, STREAM_XSTAR, cj) < w_density
    w_star = np.where(mask, normal01(seed, STREAM_ALPHA, cj) * 8.0, 0.0)
    logits = np.zeros(m)
    np.add.at(logits, np.repeat(np.arange(m), np.diff(rowptr)), values * w_star[colind])
    prob = 1.0 / (1.0 + np.exp(-(logits - 0.1)))
    y = (uniform01(seed, STREAM_YSTAR, rows) < prob).astype(np.float64)
    return rowptr, colind, values, y


def logreg_lambda_max(X, y) -> float:
    """Smallest l1 weight for which w = 0 (intercept 0) is stationary for the feature weights of the logistic loss of
    sparse_logreg/runme.jl:18-39: |X'(1/2 - y)|_inf / N.  The reference uses lam = 0.01 on data sets whose lambda_max is a
    few tenths (sparse_logreg/runme.jl:182); configs scale lam as a fraction of this value so the instance is not degenerate."""
    y = np.asarray(y, dtype=np.float64)
    return float(np.max(np.abs(X.T @ (0.5 - y))) / y.shape[0])


def dense_classification(m=50000, n=2000, seed=0, flip=0.10):
    """X ~ N(0,1)/sqrt(n) (C order) and labels +-1 from a planted hyperplane
    with a fraction ``flip`` of the labels flipped (config C3, dual SVM)."""
    i = np.arange(m, dtype=np.uint64)[:, None]
    j = np.arange(n, dtype=np.uint64)[None, :]
    X = normal01(seed, STREAM_MATRIX, i * np.uint64(n) + j) / np.sqrt(n)
    w = normal01(seed, STREAM_XSTAR, np.arange(n, dtype=np.uint64))
    s = np.sign(X @ w)
    s[s == 0] = 1.0
    fl = uniform01(seed, STREAM_YSTAR, np.arange(m, dtype=np.uint64)) < flip
    y = np.where(fl, -s, s)
    return X, y


def dense_regression(m=50000, n=2000, seed=0, noise=0.1):
    """X ~ N(0,1)/sqrt(n) and targets X w* + Laplace noise (config C3, LAD and
    square-root lasso).  Returns (X, y)."""
    i = np.arange(m, dtype=np.uint64)[:, None]
    j = np.arange(n, dtype=np.uint64)[None, :]
    X = normal01(seed, STREAM_MATRIX, i * np.uint64(n) + j) / np.sqrt(n)
    cj = np.arange(n, dtype=np.uint64)
    w = np.where(uniform01(seed, STREAM_XSTAR, cj) < 0.05, normal01(seed, STREAM_ALPHA, cj) * 3.0, 0.0)
    u = uniform01(seed, STREAM_YSTAR, np.arange(m, dtype=np.uint64)) - 0.5
    lap = -noise * np.sign(u) * np.log(1.0 - 2.0 * np.abs(u))
    return X, X @ w + lap
