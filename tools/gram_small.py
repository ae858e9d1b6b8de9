"""One short dual-SVM solve in the Gram form at the BASELINE configs[2] size (for profiler captures of k_primal_dual):
N = 50000, d = 2000, AdaPDM/OurRule, 5 iterations + prologue = 6 evaluations = 12 sweeps over Z (0.8 GB each)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import adaprox_b200 as AdaProx  # noqa: E402
dev = AdaProx.Device(0); AdaProx.set_default_device(dev)
N, d = 50000, 2000
X, y = AdaProx.synth.dense_classification(N, d, 0)
f = AdaProx.QuadraticGram(y[:, None] * X, -np.ones(N))
A = AdaProx.DeviceMatrix(y[None, :].copy())
x, yy, it = AdaProx.adaptive_primal_dual(np.zeros(N), np.zeros(1), f=f, g=AdaProx.IndBox(0.0, 0.1), h=AdaProx.IndZero(), A=A,
                                         rule=AdaProx.OurRule(t=0.1, norm_A=float(np.sqrt(N))), tol=0.0, maxit=5)
print("ok", it, AdaProx.last_solve_info()["solve_ms"])
