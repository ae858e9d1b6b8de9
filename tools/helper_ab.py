"""A/B of the helper CTAs beside the single-sweep kernel (solver_fused_helper.cuh) on the headline instance: ms per iteration with
ADAPROX_HELPERS = 0 / 2 and several batch sizes, and a bitwise comparison of the iterates (the result must not depend on who
processed which chunk).   python tools/helper_ab.py [m n]"""
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, ".")
import adaprox_b200 as AdaProx  # noqa: E402

m, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (65536, 131072)
P = AdaProx.generate_planted_lasso(m, n, pfactor=5, seed=0, power_iters=2)
f, g = AdaProx.LinearLeastSquares(P["A"], P["b"]), AdaProx.NormL1(1.0)
ref = None
for helpers, rows in ((0, 8), (2, 4), (2, 8), (2, 16), (2, 32), (0, 8), (2, 8), (0, 8)):
    os.environ["ADAPROX_HELPERS"] = str(helpers)
    os.environ["ADAPROX_HELPER_ROWS"] = str(rows)
    log = []
    x, it = AdaProx.adaptive_proxgrad(np.zeros(n), f=f, g=g, rule=AdaProx.OurRule(gamma=1 / P["Lf"]), tol=0.0, maxit=40, log=log)
    info = AdaProx.last_solve_info()
    h = hashlib.sha1(x.tobytes()).hexdigest()[:12]
    ref = ref or h
    print(json.dumps(dict(helpers=helpers, batch_rows=rows, ms_per_iteration=info["solve_ms"] / 40, its_per_s=40e3 / info["solve_ms"], launches=info["kernel_launches"],
                          x_sha1=h, same_bits_as_first=(h == ref), gamma40=log[-1]["gamma"], norm_res=log[-1]["norm_res"])), flush=True)
