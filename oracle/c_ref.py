"""ctypes loader of oracle/libadaprox_ref.so (the plain-C restatement, oracle/adaprox_ref.c).  TEST INFRASTRUCTURE:
imported by tests/ only.  Builds the library with gcc on first use if it is missing."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
F_ZERO, F_LEAST_SQUARES, F_LOGISTIC, F_QUADRATIC = range(4)
P_ZERO, P_IND_ZERO, P_NORM_L1, P_NORM_L2, P_IND_BOX = range(5)
RULE_FIXED, RULE_MM, RULE_OUR = range(3)
_dp = C.POINTER(C.c_double)


class Prox(C.Structure):
    _fields_ = [("kind", C.c_int), ("lam", C.c_double), ("lo", C.c_double), ("hi", C.c_double), ("shift", _dp)]


class Problem(C.Structure):
    _fields_ = [("f_kind", C.c_int), ("F", _dp), ("fm", C.c_long), ("fn", C.c_long), ("fvec", _dp), ("g", Prox), ("h", Prox),
                ("A", _dp), ("am", C.c_long), ("n", C.c_long), ("rule", C.c_int), ("gamma", C.c_double), ("t", C.c_double),
                ("norm_A", C.c_double), ("delta", C.c_double), ("Theta", C.c_double), ("tol", C.c_double), ("maxit", C.c_long)]


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


def adaptive_primal_dual(x0, y0, *, f_kind, F=None, fvec=None, g, h=None, A=None, rule, gamma, t=1.0, norm_A=0.0, delta=0.0,
                         Theta=1.2, tol=1e-5, maxit=10_000, nhist=0):
    """src/AdaProx.jl:312-364 through the C restatement; A=None is `adaptive_proxgrad` (:418-421).
    g, h: (Prox, keepalive) pairs from prox_desc.  Returns (x, y, it, hist) with hist = dict of gamma/sigma/norm_res/objective."""
    lib = load()
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    n = x0.shape[0]
    Ff = None if F is None else np.asfortranarray(F, dtype=np.float64)
    Af = None if A is None else np.asfortranarray(A, dtype=np.float64)
    fv = None if fvec is None else np.ascontiguousarray(fvec, dtype=np.float64)
    md = Af.shape[0] if Af is not None else n
    y0 = np.zeros(md) if y0 is None else np.ascontiguousarray(y0, dtype=np.float64)
    if h is None:
        h = prox_desc(P_ZERO)
    p = Problem(f_kind, _p(Ff), Ff.shape[0] if Ff is not None else 0, Ff.shape[1] if Ff is not None else 0, _p(fv), g[0], h[0],
                _p(Af), md if Af is not None else 0, n, rule, float(gamma), float(t), float(norm_A), float(delta), float(Theta),
                float(tol), int(maxit))
    x, y = np.empty(n), np.empty(md)
    H = int(min(nhist, maxit))
    hist = {k: np.empty(max(H, 1)) for k in ("gamma", "sigma", "norm_res", "objective")}
    it = lib.ref_adaptive_primal_dual(C.byref(p), _p(x0), _p(y0), _p(x), _p(y), _p(hist["gamma"]), _p(hist["sigma"]),
                                      _p(hist["norm_res"]), _p(hist["objective"]), H)
    if it < 0:
        raise MemoryError("ref_adaptive_primal_dual: allocation failed")
    k = min(H, it)
    return x, y, int(it), {kk: vv[:k] for kk, vv in hist.items()}
