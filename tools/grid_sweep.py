"""A/B of the persistent-kernel grid size (ADAPROX_GRID) on the latency-bound configs: C1 lasso 400x1000, C2 sparse
logreg (rcv1 shape), a mid-size dense lasso and the C3 LAD instance.  One JSON line per (config, grid).
The environment variable is read at every adaprox_solve, so one process sweeps all grids."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, ".")
import adaprox_b200 as AdaProx  # noqa: E402


def run(name, fn, grids, reps=3):
    for G in grids:
        os.environ["ADAPROX_GRID"] = str(G)
        fn()
        best, it = None, None
        for _ in range(reps):
            it = fn()
            ms = AdaProx.last_solve_info()["solve_ms"]
            best = ms if best is None else min(best, ms)
        print(json.dumps(dict(config=name, grid=G, iterations=it, device_ms=best, us_per_iteration=1e3 * best / it)), flush=True)
    os.environ.pop("ADAPROX_GRID", None)


def main():
    which = sys.argv[1:] or ["c1", "c2", "mid", "lad"]
    grids = [296, 222, 148, 111, 74, 37, 16]
    AdaProx.default_device()
    if "c1" in which:
        P = AdaProx.synth.planted_lasso(400, 1000, 5, 0)
        Lf = AdaProx.synth.spectral_norm_sq(P["A"], iters=1000, tol=1e-15)
        f, g = AdaProx.LinearLeastSquares(P["A"], P["b"]), AdaProx.NormL1(1.0)
        run("C1 lasso 400x1000 AdaPGM", lambda: AdaProx.adaptive_proxgrad(np.zeros(1000), f=f, g=g, rule=AdaProx.OurRule(gamma=1 / Lf), tol=1e-6, maxit=10000)[1], grids)
    if "c2" in which:
        import scipy.sparse as sp
        rp, ci, va, y = AdaProx.synth.sparse_logreg(20242, 47236, 0)
        X = sp.csr_matrix((va, ci, rp), shape=(20242, 47236))
        n = 47237
        gam = 4 * 20242 / (va @ va + 20242)
        f, g = AdaProx.LogisticLoss(X, y), AdaProx.NormL1(1e-4)
        run("C2 sparse logreg 20242x47236 AdaPGM (200 iterations)", lambda: AdaProx.adaptive_proxgrad(np.zeros(n), f=f, g=g, rule=AdaProx.OurRule(gamma=gam), tol=0.0, maxit=200)[1], grids)
    if "mid" in which:
        for m, n in ((2048, 4096), (8192, 8192)):
            P = AdaProx.synth.planted_lasso(m, n, 5, 0)
            Lf = AdaProx.synth.spectral_norm_sq(P["A"], iters=30)
            f, g = AdaProx.LinearLeastSquares(P["A"], P["b"]), AdaProx.NormL1(1.0)
            run(f"lasso {m}x{n} AdaPGM (300 iterations)", lambda: AdaProx.adaptive_proxgrad(np.zeros(n), f=f, g=g, rule=AdaProx.OurRule(gamma=1 / Lf), tol=0.0, maxit=300)[1], grids)
    if "paper" in which:
        # shapes of the reference's own experiment datasets (experiments/*/runme.jl main()): mushrooms 8124 x 112 with 21 nonzeros
        # per row (sparse logreg), cpusmall 8192 x 12 (+ intercept column; LAD with AdaPDM+), svmguide3 1243 x 22 (dual SVM, dense Q)
        import scipy.sparse as sp
        rng = np.random.default_rng(0)
        m, n, k = 8124, 112, 21
        cols = np.concatenate([rng.choice(n, k, replace=False) for _ in range(m)])
        X = sp.csr_matrix((np.ones(m * k), cols, np.arange(0, m * k + 1, k)), shape=(m, n))
        w = rng.standard_normal(n)
        y = (X @ w + 0.5 * rng.standard_normal(m) > 0).astype(float)
        f, g = AdaProx.LogisticLoss(X, y), AdaProx.NormL1(1e-3)
        gam = 4 * m / (X.nnz + m)
        run("mushrooms-shaped sparse logreg 8124x112 AdaPGM (300 iterations)", lambda: AdaProx.adaptive_proxgrad(np.zeros(n + 1), f=f, g=g, rule=AdaProx.OurRule(gamma=gam), tol=0.0, maxit=300)[1], grids)
        A = np.hstack([rng.standard_normal((8192, 12)), np.ones((8192, 1))])
        yv = A @ rng.standard_normal(13) + rng.laplace(size=8192)
        nA = float(np.linalg.norm(A))
        Ad = AdaProx.DeviceMatrix(A)
        kw = dict(f=AdaProx.Zero(), g=AdaProx.NormL1(1.0), h=AdaProx.Translate(AdaProx.NormL1(), -yv), A=Ad, eta=nA, t=1.0, tol=0.0, maxit=300)
        run("cpusmall-shaped LAD 8192x13 AdaPDM+ (300 iterations)", lambda: AdaProx.adaptive_linesearch_primal_dual(np.zeros(13), np.zeros(8192), **kw)[2], grids)
        Xs, ys = AdaProx.synth.dense_classification(1243, 22, 0)
        Z = ys[:, None] * Xs
        fq = AdaProx.Quadratic(Z @ Z.T, -np.ones(1243))
        Am = AdaProx.DeviceMatrix(ys[None, :].copy())
        run("svmguide3-shaped dual SVM N=1243 dense Q AdaPDM (300 iterations)", lambda: AdaProx.adaptive_primal_dual(np.zeros(1243), np.zeros(1), f=fq, g=AdaProx.IndBox(0.0, 0.1), h=AdaProx.IndZero(), A=Am, rule=AdaProx.OurRule(t=1.0, norm_A=float(np.sqrt(1243))), tol=0.0, maxit=300)[2], grids)
    if "lad" in which:
        X, yv = AdaProx.synth.dense_regression(50000, 2000, 0)
        A = np.hstack([X, np.ones((50000, 1))])
        nA = float(np.linalg.norm(A))
        Ad = AdaProx.DeviceMatrix(A)
        h = AdaProx.Translate(AdaProx.NormL1(), -yv)
        kw = dict(f=AdaProx.Zero(), g=AdaProx.NormL1(10.0), h=h, A=Ad, eta=nA, t=1.0, tol=0.0, maxit=300)
        run("C3 lad 50000x2001 AdaPDM+ (300 iterations)", lambda: AdaProx.adaptive_linesearch_primal_dual(np.zeros(2001), np.zeros(50000), **kw)[2], [296, 148])


if __name__ == "__main__":
    main()
