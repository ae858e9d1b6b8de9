"""One short batched lambda-path solve at the config-5 size (for profiler captures of k_path_gemm)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import adaprox_b200 as AdaProx  # noqa: E402
dev = AdaProx.Device(0); AdaProx.set_default_device(dev)
m, n, Lc = 16384, 8192, 256
P = AdaProx.generate_planted_lasso(m, n, pfactor=5, seed=0, power_iters=2)
f = AdaProx.LinearLeastSquares(P["A"], P["b"])
lambdas = np.linspace(0.5, 5.0, Lc)
X, its, info = AdaProx.adaptive_proxgrad_path(None, f=f, lambdas=lambdas, rule=AdaProx.OurRule(gamma=1 / P["Lf"]), tol=0.0, maxit=3)
print("ok", info["solve_ms"], info["batched_evals"])
