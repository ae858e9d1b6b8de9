"""Small single-pass fused AdaPGM solve (forced on with ADAPROX_FUSED=1) -- used to probe profiler compatibility."""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
os.environ.setdefault("ADAPROX_FUSED", "1")
import adaprox_b200 as AdaProx  # noqa: E402

m, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (96, 20000)
P = AdaProx.synth.planted_lasso(m, n, 60, 4)
f = AdaProx.LinearLeastSquares(P["A"], P["b"])
x, it = AdaProx.adaptive_proxgrad(np.zeros(n), f=f, g=AdaProx.NormL1(1.0), rule=AdaProx.OurRule(gamma=1e-3), tol=0.0, maxit=5)
print("ok", it, AdaProx.last_solve_info())
