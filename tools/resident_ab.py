import sys, os, time, json
sys.path.insert(0, '/root/repo')
import numpy as np
import adaprox_b200 as AdaProx
from oracle import adaprox_oracle as O
P = AdaProx.synth.planted_lasso(400, 1000, 5, 0)
Lf = AdaProx.synth.spectral_norm_sq(P["A"], iters=1000, tol=1e-15)
f, g = AdaProx.LinearLeastSquares(P["A"], P["b"]), AdaProx.NormL1(1.0)
logo=[]; xo, ito = O.adaptive_proxgrad(np.zeros(1000), f=O.LinearLeastSquares(P["A"], P["b"]), g=O.NormL1(1.0), rule=O.OurRule(gamma=1/Lf), tol=1e-6, maxit=10000, log=logo)
for mode in ("1", "0"):
    os.environ["ADAPROX_RESIDENT"] = mode
    log=[]
    AdaProx.adaptive_proxgrad(np.zeros(1000), f=f, g=g, rule=AdaProx.OurRule(gamma=1/Lf), tol=1e-6, maxit=100)
    x, it = AdaProx.adaptive_proxgrad(np.zeros(1000), f=f, g=g, rule=AdaProx.OurRule(gamma=1/Lf), tol=1e-6, maxit=10000, log=log)
    info = AdaProx.last_solve_info()
    gd=np.array([r["gamma"] for r in log[:40]]); go=np.array([r["gamma"] for r in logo[:40]])
    print(json.dumps(dict(resident=mode, it=it, oracle_it=ito, us_per_it=1e3*info["solve_ms"]/it, passes=info["matrix_passes"], gam15=float(np.max(np.abs(gd[:15]/go[:15]-1))), gam40=float(np.max(np.abs(gd/go-1))),
          obj=float(abs(log[-1]["objective"]-logo[-1]["objective"])/logo[-1]["objective"]), xerr=float(np.linalg.norm(x-xo)/np.linalg.norm(xo)))))
    x2, it2 = AdaProx.adaptive_proxgrad(np.zeros(1000), f=f, g=g, rule=AdaProx.OurRule(gamma=1/Lf), tol=1e-6, maxit=10000)
    print("no-log run:", it2, 1e3*AdaProx.last_solve_info()["solve_ms"]/it2, "us/it; bit-identical:", np.array_equal(x, x2))
