"""A/B of the CSR sweep layout (lanes per row, ADAPROX_CSR_LPR; read at matrix upload) on the rcv1-shaped sparse logistic
regression (BASELINE configs[1]): per-iteration time of AdaPGM over 200 iterations and of the two operator calls alone.
One JSON line per setting; "auto" is what the library picks from the mean row length."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, ".")
import adaprox_b200 as AdaProx  # noqa: E402


def main():
    import scipy.sparse as sp
    AdaProx.default_device()
    rp, ci, va, y = AdaProx.synth.sparse_logreg(20242, 47236, 0)
    X = sp.csr_matrix((va, ci, rp), shape=(20242, 47236))
    n = 47237
    gam = 4 * 20242 / (va @ va + 20242)
    g = AdaProx.NormL1(1e-4)
    ref = None
    for lpr in ("32", "16", "8", "4", "auto"):
        if lpr == "auto":
            os.environ.pop("ADAPROX_CSR_LPR", None)
        else:
            os.environ["ADAPROX_CSR_LPR"] = lpr
        M = AdaProx.DeviceMatrix(X)
        f = AdaProx.LogisticLoss(M, y)
        best = None
        for _ in range(4):
            x, it = AdaProx.adaptive_proxgrad(np.zeros(n), f=f, g=g, rule=AdaProx.OurRule(gamma=gam), tol=0.0, maxit=200)
            ms = AdaProx.last_solve_info()["solve_ms"]
            best = ms if best is None else min(best, ms)
        if os.environ.get("CSR_SWEEP_PHASES"):      # per-phase breakdown of the same solve (library prints it to stderr)
            print(f"--- lanes_per_row={lpr}", file=sys.stderr, flush=True)
            os.environ["ADAPROX_PHASE_TIMING"] = "1"
            AdaProx.adaptive_proxgrad(np.zeros(n), f=f, g=g, rule=AdaProx.OurRule(gamma=gam), tol=0.0, maxit=60)
            os.environ.pop("ADAPROX_PHASE_TIMING")
        obj = float(f(x) + g(x))
        ref = obj if ref is None else ref
        print(json.dumps(dict(config="C2 sparse logreg 20242x47236 AdaPGM (200 iterations)", lanes_per_row=lpr, nnz=int(len(va)),
                              us_per_iteration=1e3 * best / it, objective=obj, rel_diff_vs_first=abs(obj - ref) / abs(ref))), flush=True)
        M.free()


if __name__ == "__main__":
    main()
