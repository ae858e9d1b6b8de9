"""Determinism probe: the fused solve repeated on a small ragged instance must give bit-identical stepsizes."""
import os, sys
import numpy as np
sys.path.insert(0, ".")
os.environ["ADAPROX_FUSED"] = "1"
import adaprox_b200 as AdaProx  # noqa: E402
from oracle import adaprox_oracle as O  # noqa: E402
m, n, pf = 96, 9000, 60
P = AdaProx.synth.planted_lasso(m, n, pf, 4)
Lf = AdaProx.synth.spectral_norm_sq(P["A"], iters=300)
logo = []
O.adaptive_proxgrad(np.zeros(n), f=O.LinearLeastSquares(P["A"], P["b"]), g=O.NormL1(1.0), rule=O.OurRule(gamma=1 / Lf), tol=1e-6, maxit=60, log=logo)
go = np.array([r["gamma"] for r in logo[:25]])
f = AdaProx.LinearLeastSquares(P["A"], P["b"])
ref = None
for rep in range(12):
    log = []
    AdaProx.adaptive_proxgrad(np.zeros(n), f=f, g=AdaProx.NormL1(1.0), rule=AdaProx.OurRule(gamma=1 / Lf), tol=1e-6, maxit=60, log=log)
    gd = np.array([r["gamma"] for r in log[:25]])
    if ref is None: ref = gd
    print(rep, "max rel dev vs oracle (first 12 / 25): %.2e %.2e" % (np.max(np.abs(gd[:12] / go[:12] - 1)), np.max(np.abs(gd / go - 1))),
          "bit-identical to run 0:", bool(np.array_equal(gd, ref)))
