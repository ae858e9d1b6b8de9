"""Rounding-drift envelopes (TEST INFRASTRUCTURE, like everything under oracle/).

The AdaPGM / AdaPDM stepsize recursion amplifies rounding differences (src/AdaProx.jl:258-273: gamma depends on
ratios of differences of nearly equal vectors), so two Float64 evaluations of the same algorithm that only differ in
summation order separate exponentially.  `north_star` asks for stepsizes within 1e-12 of the
reference; that is only meaningful while Float64 itself determines the trajectory to 1e-12.  This module measures how
far that is: the oracle is run (a) in Float64, (b) in Float64 with the columns of the matrix permuted (another valid
summation order, identical in exact arithmetic) and (c) in x87 extended precision (`adaprox_oracle.precision`); the
distance of (a) and (b) from (c) is the INTRINSIC drift.  The device path passes when its own distance from (c) stays
below `factor` x that envelope (never tighter than `floor`).
"""
from __future__ import annotations

import numpy as np

from . import adaprox_oracle as O

LD = np.longdouble


def series(log, key="gamma"):
    return np.array([r[key] for r in log], dtype=LD)


def collect(run, nperm=2, seed=0):
    """Generic form: ``run(dtype, rng)`` returns the record list of one oracle run -- ``dtype`` is np.longdouble inside
    ``O.precision`` (build arrays, rules and prox objects with it) or np.float64; ``rng`` is None for the natural order or a
    Generator the callee uses to permute its data (rows / columns: the same problem in exact arithmetic)."""
    out = {}
    with O.precision(LD):
        out["ext"] = run(LD, None)
    out["f64"] = run(np.float64, None)
    rng = np.random.default_rng(seed)
    out["perms"] = [run(np.float64, rng) for _ in range(nperm)]
    return out


def lasso_runs(A, b, lam, make_rule, K, nperm=2, seed=0, tol=0.0):
    """AdaPGM on 1/2|Ax-b|^2 + lam|x|_1 for K iterations: returns dict(ext=log, f64=log, perms=[log...])."""
    A = np.asarray(A, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    n = A.shape[1]
    out = {}
    with O.precision(LD):
        log = []
        O.adaptive_proxgrad(np.zeros(n, dtype=LD), f=O.LinearLeastSquares(A.astype(LD), b.astype(LD)), g=O.NormL1(lam),
                            rule=make_rule(O), tol=tol, maxit=K, log=log)
        out["ext"] = log
    log = []
    O.adaptive_proxgrad(np.zeros(n), f=O.LinearLeastSquares(np.asfortranarray(A), b), g=O.NormL1(lam), rule=make_rule(O), tol=tol, maxit=K, log=log)
    out["f64"] = log
    out["perms"] = []
    rng = np.random.default_rng(seed)
    for _ in range(nperm):
        perm = rng.permutation(n)
        log = []
        O.adaptive_proxgrad(np.zeros(n), f=O.LinearLeastSquares(np.asfortranarray(A[:, perm]), b), g=O.NormL1(lam), rule=make_rule(O), tol=tol,
                            maxit=K, log=log)
        out["perms"].append(log)
    return out


def rel_to(ref, x):
    ref = np.asarray(ref, dtype=LD)
    x = np.asarray(x, dtype=LD)
    k = min(len(ref), len(x))
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.abs(x[:k] / ref[:k] - 1).astype(np.float64)


def envelope(runs, key="gamma"):
    """Running maximum over iterations of the largest Float64-oracle distance from the extended-precision run."""
    ext = series(runs["ext"], key)
    d = [rel_to(ext, series(runs["f64"], key))] + [rel_to(ext, series(p, key)) for p in runs["perms"]]
    k = min(len(x) for x in d)
    return np.maximum.accumulate(np.max(np.stack([x[:k] for x in d]), axis=0))


def check_inside(device_series, runs, key="gamma", factor=20.0, floor=1e-12):
    """-> (ok, device_drift, allowed) with allowed[k] = max(floor, factor * envelope[k])."""
    ext = series(runs["ext"], key)
    dd = rel_to(ext, np.asarray(device_series))
    env = envelope(runs, key)
    k = min(len(dd), len(env))
    allowed = np.maximum(floor, factor * env[:k])
    return bool(np.all(dd[:k] <= allowed)), dd[:k], allowed


def perm_envelope(base, perms):
    """Float64-only envelope for problems too large for an extended-precision run: running maximum over the iterations of the
    largest relative distance between the oracle's natural-order run `base` and its permuted-data runs `perms` (same algorithm,
    other summation orders).  One permutation is a single sample of a heavy-tailed ratio (the stepsize rule divides differences
    of nearly equal numbers, src/AdaProx.jl:260-268); use three or more."""
    base = np.asarray(base, dtype=np.float64)
    d = [np.abs(np.asarray(p, dtype=np.float64)[: len(base)] / base[: len(p)] - 1) for p in perms]
    k = min(len(x) for x in d)
    return np.maximum.accumulate(np.max(np.stack([x[:k] for x in d]), axis=0))
