"""A/B of the lambda-path DMMA contractions (config 5 shape): ms and TFLOP/s per contraction and per batched iteration.
Usage: [ADAPROX_LIB=build_ab/libadaprox_<variant>.so] python tools/path_gemm_ab.py"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, ".")
import adaprox_b200 as AdaProx  # noqa: E402

m, n = 16384, 8192
P = AdaProx.generate_planted_lasso(m, n, pfactor=5, seed=0, power_iters=5)
f = AdaProx.LinearLeastSquares(P["A"], P["b"])
for Lc in (256, 32):
    ms_r = P["A"].time_path_gemm(Lc, 0, reps=10)
    ms_g = P["A"].time_path_gemm(Lc, 1, reps=10)
    lam = np.linspace(1.0, 0.01, Lc)
    AdaProx.adaptive_proxgrad_path(None, f=f, lambdas=lam, rule=AdaProx.OurRule(gamma=1 / P["Lf"]), tol=0.0, maxit=3)
    X, its, info = AdaProx.adaptive_proxgrad_path(None, f=f, lambdas=lam, rule=AdaProx.OurRule(gamma=1 / P["Lf"]), tol=0.0, maxit=30)
    ev = info["batched_evals"]
    print(json.dumps(dict(lib=os.environ.get("ADAPROX_LIB", "default"), L=Lc, AX_ms=ms_r, AtR_ms=ms_g, AX_tflops=2.0 * m * n * Lc / (ms_r * 1e-3) / 1e12,
                          AtR_tflops=2.0 * m * n * Lc / (ms_g * 1e-3) / 1e12, ms_per_iteration=info["solve_ms"] / ev,
                          tflops_sustained=4.0 * m * n * Lc * ev / (info["solve_ms"] * 1e-3) / 1e12, checksum=float(np.abs(X).sum()))), flush=True)
