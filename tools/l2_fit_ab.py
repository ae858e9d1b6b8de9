"""Two-pass dense AdaPGM on matrices around the L2 size (126 MB): does loading the stream evict_first (shipped) cost anything
when the whole matrix could have stayed in L2?  Run with ADAPROX_L2_KEEP_MB=0 (shipped) and -1 (no hints)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, ".")
os.environ.setdefault("ADAPROX_FUSED", "0")           # ADAPROX_FUSED=1: the single-sweep kernel on the same shapes (where is the crossover?)
os.environ.setdefault("ADAPROX_RESIDENT", "0")
import adaprox_b200 as AdaProx  # noqa: E402

AdaProx.default_device()
SHAPES = os.environ.get("SHAPES")
shapes = [tuple(int(v) for v in t.split(",")) for t in SHAPES.split(";")] if SHAPES else \
    [(1000, 4096), (2000, 4096), (3000, 4096), (4000, 4096), (6000, 4096), (8000, 8192)]
for m, n in shapes:
    P = AdaProx.generate_planted_lasso(m, n, pfactor=5, seed=0, power_iters=30)
    f, g = AdaProx.LinearLeastSquares(P["A"], P["b"]), AdaProx.NormL1(1.0)
    out = []
    for rep in range(3):
        x, it = AdaProx.adaptive_proxgrad(np.zeros(n), f=f, g=g, rule=AdaProx.OurRule(gamma=1 / P["Lf"]), tol=0.0, maxit=400)
        info = AdaProx.last_solve_info()
        out.append(1e3 * info["solve_ms"] / it)
    print(json.dumps(dict(m=m, n=n, MB=round(m * n * 8 / 1e6), keep=os.environ.get("ADAPROX_L2_KEEP_MB", "0"), fused=os.environ["ADAPROX_FUSED"],
                          us_per_iteration=[round(v, 2) for v in out], passes=info.get("matrix_passes"))), flush=True)
