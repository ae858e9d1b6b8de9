"""ctypes loader of oracle/libadaprox_ref.so (the plain-C restatement, oracle/adaprox_ref.c).  TEST INFRASTRUCTURE:
imported by tests/ only.  Builds the library with gcc on first use if it is missing."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
F_ZERO, F_LEAST_SQUARES, F_LOGISTIC, F_QUADRATIC = range(4)
P_ZERO, P_IND_ZERO, P_NORM_L1, P_NORM_L2, P_IND_BOX = range(5)
RULE_FIXED, RULE_MM, RULE_OUR, RULE_OUR_PLUS = range(4)
_dp = C.POINTER(C.c_double)


class Prox(C.Structure):
    _fields_ = [("kind", C.c_int), ("lam", C.c_double), ("lo", C.c_double), ("hi", C.c_double), ("shift", _dp)]


class Problem(C.Structure):
    _fields_ = [("f_kind", C.c_int), ("F", _dp), ("fm", C.c_long), ("fn", C.c_long), ("fvec", _dp), ("g", Prox), ("h", Prox),
                ("A", _dp), ("am", C.c_long), ("n", C.c_long), ("rule", C.c_int), ("gamma", C.c_double), ("t", C.c_double),
                ("norm_A", C.c_double), ("delta", C.c_double), ("Theta", C.c_double), ("tol", C.c_double), ("maxit", C.c_long),
                ("xi", C.c_double), ("nu", C.c_double), ("r", C.c_double)]


_lib = None


def load():
    global _lib
    if _lib is None:
        so = os.path.join(HERE, "libadaprox_ref.so")
        src = os.path.join(HERE, "adaprox_ref.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.run(["make", "-C", HERE], check=True, capture_output=True)
        _lib = C.CDLL(so)
        _lib.ref_adaptive_primal_dual.restype = C.c_long
        _lib.ref_adaptive_primal_dual.argtypes = [C.POINTER(Problem), _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, C.c_long]
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def prox_desc(kind, lam=1.0, lo=0.0, hi=0.0, shift=None):
    """returns (Prox, keepalive)"""
    s = None if shift is None else np.ascontiguousarray(shift, dtype=np.float64)
    return Prox(kind, float(lam), float(lo), float(hi), _p(s)), s


def _problem(x0, *, f_kind, F, fvec, g, h, A, rule=RULE_FIXED, gamma=0.0, t=1.0, norm_A=0.0, delta=0.0, Theta=1.2, tol=1e-5, maxit=10_000,
             xi=1.0, nu=1.0, r=0.5):
    """builds the C problem struct; returns (Problem, x0, n, md, keepalive)"""
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    n = x0.shape[0]
    Ff = None if F is None else np.asfortranarray(F, dtype=np.float64)
    Af = None if A is None else np.asfortranarray(A, dtype=np.float64)
    fv = None if fvec is None else np.ascontiguousarray(fvec, dtype=np.float64)
    md = Af.shape[0] if Af is not None else n
    if h is None:
        h = prox_desc(P_ZERO)
    p = Problem(f_kind, _p(Ff), Ff.shape[0] if Ff is not None else 0, Ff.shape[1] if Ff is not None else 0, _p(fv), g[0], h[0],
                _p(Af), md if Af is not None else 0, n, rule, float(gamma), float(t), float(norm_A), float(delta), float(Theta),
                float(tol), int(maxit), float(xi), float(nu), float(r))
    return p, x0, n, md, (Ff, Af, fv, g, h)


def _hist(nhist, maxit, keys):
    H = int(min(nhist, maxit))
    return H, {k: np.empty(max(H, 1)) for k in keys}


def adaptive_primal_dual(x0, y0, *, f_kind, F=None, fvec=None, g, h=None, A=None, rule, gamma, t=1.0, norm_A=0.0, delta=0.0,
                         Theta=1.2, tol=1e-5, maxit=10_000, nhist=0, xi=1.0, nu=1.0, r=0.5):
    """src/AdaProx.jl:312-364 through the C restatement; A=None is `adaptive_proxgrad` (:418-421).
    g, h: (Prox, keepalive) pairs from prox_desc.  Returns (x, y, it, hist) with hist = dict of gamma/sigma/norm_res/objective."""
    lib = load()
    p, x0, n, md, keep = _problem(x0, f_kind=f_kind, F=F, fvec=fvec, g=g, h=h, A=A, rule=rule, gamma=gamma, t=t, norm_A=norm_A,
                                  delta=delta, Theta=Theta, tol=tol, maxit=maxit, xi=xi, nu=nu, r=r)
    y0 = np.zeros(md) if y0 is None else np.ascontiguousarray(y0, dtype=np.float64)
    x, y = np.empty(n), np.empty(md)
    H, hist = _hist(nhist, maxit, ("gamma", "sigma", "norm_res", "objective"))
    it = lib.ref_adaptive_primal_dual(C.byref(p), _p(x0), _p(y0), _p(x), _p(y), _p(hist["gamma"]), _p(hist["sigma"]),
                                      _p(hist["norm_res"]), _p(hist["objective"]), H)
    if it < 0:
        raise MemoryError("ref_adaptive_primal_dual: allocation failed")
    k = min(H, it)
    return x, y, int(it), {kk: vv[:k] for kk, vv in hist.items()}


def adaptive_linesearch_primal_dual(x0, y0, *, f_kind, F=None, fvec=None, g, h, A, gamma=None, eta=1.0, t=1.0, delta=1e-8, Theta=1.2,
                                    r=2.0, R=0.95, tol=1e-5, maxit=10_000, nhist=0):
    """src/AdaProx.jl:463-550 (AdaPDM+).  Returns (x, y, it, hist, trials)."""
    lib = load()
    if gamma is None:
        gamma = 1.0 / (2.0 * Theta * t * eta)                                     # :484-486
    p, x0, n, md, keep = _problem(x0, f_kind=f_kind, F=F, fvec=fvec, g=g, h=h, A=A, gamma=gamma, t=t, delta=delta, Theta=Theta,
                                  tol=tol, maxit=maxit)
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    x, y = np.empty(n), np.empty(md)
    H, hist = _hist(nhist, maxit, ("gamma", "sigma", "norm_res", "objective"))
    trials = C.c_long(0)
    lib.ref_adaptive_linesearch_primal_dual.restype = C.c_long
    lib.ref_adaptive_linesearch_primal_dual.argtypes = [C.POINTER(Problem), C.c_double, C.c_double, C.c_double, _dp, _dp, _dp, _dp,
                                                        _dp, _dp, _dp, _dp, C.c_long, C.POINTER(C.c_long)]
    it = lib.ref_adaptive_linesearch_primal_dual(C.byref(p), float(eta), float(r), float(R), _p(x0), _p(y0), _p(x), _p(y), _p(hist["gamma"]),
                                                 _p(hist["sigma"]), _p(hist["norm_res"]), _p(hist["objective"]), H, C.byref(trials))
    if it < 0:
        raise RuntimeError(f"ref_adaptive_linesearch_primal_dual failed ({it})")
    k = min(H, it)
    return x, y, int(it), {kk: vv[:k] for kk, vv in hist.items()}, int(trials.value)


BACKTRACKING_PROXGRAD, BACKTRACKING_NESTEROV, FIXED_NESTEROV, AGRAAL = range(4)


def proxgrad_family(which, x0, *, f_kind, F=None, fvec=None, g, gamma, xi=1.0, shrink=0.5, muf=0.0, mug=0.0, theta=-1.0,
                    gamma_max=1e6, phi=1.5, x_second=None, tol=1e-5, maxit=100_000, nhist=0):
    """backtracking_proxgrad (:50-64), backtracking_nesterov (:66-84), fixed_nesterov (:91-142), agraal (:150-192).
    `gamma` = gamma0 / gamma; agraal: gamma <= 0 means `gamma0 = nothing`, x_second is its other start point x0.
    Returns (x, it, hist, (f_evals, grad_evals))."""
    lib = load()
    p, x0, n, md, keep = _problem(x0, f_kind=f_kind, F=F, fvec=fvec, g=g, h=None, A=None, gamma=gamma, tol=tol, maxit=maxit)
    xs = None if x_second is None else np.ascontiguousarray(x_second, dtype=np.float64)
    x = np.empty(n)
    H, hist = _hist(nhist, maxit, ("gamma", "norm_res", "objective"))
    ev = (C.c_long * 2)()
    lib.ref_proxgrad_family.restype = C.c_long
    lib.ref_proxgrad_family.argtypes = [C.POINTER(Problem), C.c_int] + [C.c_double] * 7 + [_dp, _dp, _dp, _dp, _dp, _dp, C.c_long, C.POINTER(C.c_long)]
    it = lib.ref_proxgrad_family(C.byref(p), int(which), float(xi), float(shrink), float(muf), float(mug), float(theta), float(gamma_max),
                                 float(phi), _p(x0), _p(xs), _p(x), _p(hist["gamma"]), _p(hist["norm_res"]), _p(hist["objective"]), H, ev)
    if it < 0:
        raise MemoryError("ref_proxgrad_family: allocation failed")
    k = min(H, it)
    return x, int(it), {kk: vv[:k] for kk, vv in hist.items()}, (int(ev[0]), int(ev[1]))


def malitsky_pock(x0, y0, *, f_kind, F=None, fvec=None, g, h, A, sigma, t=1.0, tol=1e-5, maxit=10_000, nhist=0):
    """src/AdaProx.jl:581-629 with backtrack_stepsize_MP (:555-579).  Returns (x, y, it, hist)."""
    lib = load()
    p, x0, n, md, keep = _problem(x0, f_kind=f_kind, F=F, fvec=fvec, g=g, h=h, A=A, t=t, tol=tol, maxit=maxit)
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    x, y = np.empty(n), np.empty(md)
    H, hist = _hist(nhist, maxit, ("gamma", "sigma", "norm_res", "objective"))
    lib.ref_malitsky_pock.restype = C.c_long
    lib.ref_malitsky_pock.argtypes = [C.POINTER(Problem), C.c_double, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, C.c_long]
    it = lib.ref_malitsky_pock(C.byref(p), float(sigma), _p(x0), _p(y0), _p(x), _p(y), _p(hist["gamma"]), _p(hist["sigma"]),
                               _p(hist["norm_res"]), _p(hist["objective"]), H)
    if it < 0:
        raise RuntimeError(f"ref_malitsky_pock failed ({it})")
    k = min(H, it)
    return x, y, int(it), {kk: vv[:k] for kk, vv in hist.items()}


def prox_eval(desc, x, gamma, conjugate=False):
    """ProximalCore.prox(f, x, gamma) -> (y, f(y)); with conjugate=True the prox of convex_conjugate(f) (value: NaN)."""
    lib = load()
    lib.ref_prox_eval.restype = C.c_double
    lib.ref_prox_eval.argtypes = [C.POINTER(Prox), C.c_int, _dp, C.c_double, _dp, C.c_long]
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    v = lib.ref_prox_eval(C.byref(desc[0]), int(bool(conjugate)), _p(x), float(gamma), _p(y), x.shape[0])
    return y, v
