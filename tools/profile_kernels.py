"""One small workload per kernel family, for `ncu -k <kernel>` captures (profiles/r02_ncu_*.json):
    python tools/profile_kernels.py resident | gridres | c2 | pgfamily | mp | path | lad
Each prints one JSON line with the device time of the solve (NOT a bench value when run under ncu)."""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
import adaprox_b200 as AdaProx  # noqa: E402


def main():
    which = sys.argv[1]
    AdaProx.default_device()
    if which == "resident":          # k_adapgm_resident: configs[0]
        P = AdaProx.synth.planted_lasso(400, 1000, 5, 0)
        Lf = AdaProx.synth.spectral_norm_sq(P["A"], iters=1000, tol=1e-15)
        x, it = AdaProx.adaptive_proxgrad(np.zeros(1000), f=AdaProx.LinearLeastSquares(P["A"], P["b"]), g=AdaProx.NormL1(1.0),
                                          rule=AdaProx.OurRule(gamma=1 / Lf), tol=0.0, maxit=2000)
    elif which == "gridres":         # k_adapgm_gridres: the largest lasso instance of lasso/runme.jl:191-195
        P = AdaProx.synth.planted_lasso(4000, 1000, 10, 0)
        Lf = AdaProx.synth.spectral_norm_sq(P["A"], iters=100)
        x, it = AdaProx.adaptive_proxgrad(np.zeros(1000), f=AdaProx.LinearLeastSquares(P["A"], P["b"]), g=AdaProx.NormL1(1.0),
                                          rule=AdaProx.OurRule(gamma=1 / Lf), tol=0.0, maxit=2000)
        assert AdaProx.last_solve_info()["matrix_passes"] == 4
    elif which == "c2":              # k_primal_dual<false> with the CSR sweeps (spmv_rows): configs[1]
        import scipy.sparse as sp
        m, n = 20242, 47236
        rp, ci, va, y = AdaProx.synth.sparse_logreg(m, n, 0)
        X = sp.csr_matrix((va, ci, rp), shape=(m, n))
        lam = 0.03 * AdaProx.synth.logreg_lambda_max(X, y)
        x, it = AdaProx.adaptive_proxgrad(np.zeros(n + 1), f=AdaProx.LogisticLoss(X, y), g=AdaProx.NormL1(lam),
                                          rule=AdaProx.OurRule(gamma=4 * m / (va @ va + m)), tol=0.0, maxit=200)
    elif which == "pgfamily":        # k_proxgrad_family: backtracking PG on the largest lasso instance of lasso/runme.jl:192-207
        P = AdaProx.synth.planted_lasso(4000, 1000, 10, 0)
        Lf = AdaProx.synth.spectral_norm_sq(P["A"], iters=100)
        x, it = AdaProx.backtracking_proxgrad(np.zeros(1000), f=AdaProx.LinearLeastSquares(P["A"], P["b"]), g=AdaProx.NormL1(1.0),
                                              gamma0=10 / Lf, tol=0.0, maxit=300)
    elif which in ("mp", "lad"):     # k_malitsky_pock / k_primal_dual<true> on a LAD instance (least_absolute_deviation/runme.jl:39-48)
        rng = np.random.default_rng(0)
        m, d = 50000, 2000
        Xd = rng.standard_normal((m, d)) / np.sqrt(d)
        yv = Xd @ np.where(rng.random(d) < 0.05, 3.0 * rng.standard_normal(d), 0.0) + rng.laplace(scale=0.1, size=m)
        A = np.hstack([Xd, np.ones((m, 1))])
        nA = float(np.linalg.norm(A))
        h = AdaProx.Translate(AdaProx.NormL1(), -yv)
        if which == "mp":
            x, y_, it = AdaProx.malitsky_pock(np.zeros(d + 1), np.zeros(m), f=AdaProx.Zero(), g=AdaProx.NormL1(10.0), h=h, A=AdaProx.DeviceMatrix(A),
                                              sigma=1 / nA, t=1.0, tol=0.0, maxit=60)
        else:
            x, y_, it = AdaProx.adaptive_linesearch_primal_dual(np.zeros(d + 1), np.zeros(m), f=AdaProx.Zero(), g=AdaProx.NormL1(10.0), h=h,
                                                                A=AdaProx.DeviceMatrix(A), eta=nA, t=1.0, tol=0.0, maxit=60)
    elif which == "path":            # k_path_gemm<1,4> / <2,4>: configs[4]
        m, n, Lc = 16384, 8192, 256
        P = AdaProx.generate_planted_lasso(m, n, pfactor=5, seed=0, power_iters=2)
        lam = np.linspace(1.0, 0.01, Lc)
        X, its, info = AdaProx.adaptive_proxgrad_path(None, f=AdaProx.LinearLeastSquares(P["A"], P["b"]), lambdas=lam,
                                                      rule=AdaProx.OurRule(gamma=1 / P["Lf"]), tol=0.0, maxit=3)
        it = int(its.max())
    else:
        raise SystemExit(__doc__)
    info = AdaProx.last_solve_info()
    print(json.dumps(dict(workload=which, iterations=int(it), device_ms=info["solve_ms"], us_per_iteration=1e3 * info["solve_ms"] / max(int(it), 1),
                          launches=info["kernel_launches"])), flush=True)


if __name__ == "__main__":
    main()
