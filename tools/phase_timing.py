"""In-kernel phase breakdown (ADAPROX_PHASE_TIMING=1: %globaltimer stamps of CTA 0 at the phase boundaries of the persistent
primal-dual kernel) for the C3 shapes: dual SVM in the Gram form at N = 20000 and N = 50000 (d = 2000), LAD 50000 x 2001."""
import os
import sys

import numpy as np

os.environ["ADAPROX_PHASE_TIMING"] = "1"
sys.path.insert(0, ".")
import adaprox_b200 as AdaProx  # noqa: E402

AdaProx.default_device()
rng = np.random.default_rng(0)
for N in (20000, 50000):
    d = 2000
    X = rng.standard_normal((N, d)) / np.sqrt(d)
    s_ = np.sign(X @ rng.standard_normal(d)); s_[s_ == 0] = 1.0
    y = np.where(rng.random(N) < 0.1, -s_, s_)
    Zm = AdaProx.DeviceMatrix(y[:, None] * X)
    f = AdaProx.QuadraticGram(Zm, -np.ones(N))
    A = AdaProx.DeviceMatrix(y[None, :].copy())
    print(f"--- dual SVM Gram N={N}", file=sys.stderr, flush=True)
    x, yy, it = AdaProx.adaptive_primal_dual(np.zeros(N), np.zeros(1), f=f, g=AdaProx.IndBox(0.0, 0.1), h=AdaProx.IndZero(), A=A,
                                             rule=AdaProx.OurRule(t=0.1, norm_A=float(np.sqrt(N))), tol=0.0, maxit=60)
    print(f"N={N}: {1e3 * AdaProx.last_solve_info()['solve_ms'] / 60:.1f} us per iteration", file=sys.stderr, flush=True)
    Zm.free(); A.free()
m, d = 50000, 2000
Xd = rng.standard_normal((m, d)) / np.sqrt(d)
yv = Xd @ np.where(rng.random(d) < 0.05, 3.0 * rng.standard_normal(d), 0.0) + rng.laplace(scale=0.1, size=m)
Am = np.hstack([Xd, np.ones((m, 1))])
print("--- LAD 50000x2001 AdaPDM+", file=sys.stderr, flush=True)
AdaProx.adaptive_linesearch_primal_dual(np.zeros(d + 1), np.zeros(m), f=AdaProx.Zero(), g=AdaProx.NormL1(10.0), h=AdaProx.Translate(AdaProx.NormL1(), -yv),
                                        A=AdaProx.DeviceMatrix(Am), eta=float(np.linalg.norm(Am)), t=1.0, tol=0.0, maxit=60)
print(f"LAD: {1e3 * AdaProx.last_solve_info()['solve_ms'] / 60:.1f} us per iteration", file=sys.stderr, flush=True)
