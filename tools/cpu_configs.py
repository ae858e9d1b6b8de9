"""CPU arm for the other BASELINE configs: the oracle port (numpy / scipy.sparse / OpenBLAS -- the BLAS family Julia bundles) timed
on the host cores for C1 (full solve), C2 and C3 (a bounded number of iterations at full size).  One JSON line per config with
the thread count actually used.  No GPU involved; bench.py's `cpu_baseline` covers the headline config (C4)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adaprox_b200 as AdaProx  # noqa: E402  (generators only)
from oracle import adaprox_oracle as O  # noqa: E402


def threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"] or [1])
    except Exception:
        return os.cpu_count()


def line(config, it, secs, **kw):
    print(json.dumps(dict(config=config, impl="oracle port (numpy/OpenBLAS)", where=os.environ.get("CPU_CONFIGS_WHERE", "build container"),
                          cores=os.cpu_count(), blas_threads=threads(), iterations=it, seconds=secs, us_per_iteration=1e6 * secs / it,
                          iters_per_s=it / secs, **kw)), flush=True)


def main():
    which = sys.argv[1:] or ["c1", "c2", "lad", "svm"]
    if "c1" in which:
        P = AdaProx.synth.planted_lasso(400, 1000, 5, 0)
        Lf = AdaProx.synth.spectral_norm_sq(P["A"], iters=1000, tol=1e-15)
        A = np.asfortranarray(P["A"])                      # Julia's layout
        f, g = O.LinearLeastSquares(A, P["b"]), O.NormL1(1.0)
        t0 = time.perf_counter()
        x, it = O.adaptive_proxgrad(np.zeros(1000), f=f, g=g, rule=O.OurRule(gamma=1 / Lf), tol=1e-6, maxit=10000)
        line("C1 lasso 400x1000 AdaPGM OurRule tol 1e-6", it, time.perf_counter() - t0)
    if "c2" in which:
        import scipy.sparse as sp
        rp, ci, va, y = AdaProx.synth.sparse_logreg(20242, 47236, 0)
        X = sp.csc_matrix(sp.csr_matrix((va, ci, rp), shape=(20242, 47236)))       # SparseMatrixCSC, as load_libsvm_dataset returns
        gam = 4 * 20242 / (va @ va + 20242)
        f, g = O.LogisticLoss(X, y), O.NormL1(1e-4)
        O.adaptive_proxgrad(np.zeros(47237), f=f, g=g, rule=O.OurRule(gamma=gam), tol=0.0, maxit=3)
        t0 = time.perf_counter()
        x, it = O.adaptive_proxgrad(np.zeros(47237), f=f, g=g, rule=O.OurRule(gamma=gam), tol=0.0, maxit=100)
        line("C2 sparse logreg 20242x47236 AdaPGM (100 iterations, lambda 1e-4)", it, time.perf_counter() - t0, nnz=int(len(va)))
    if "lad" in which:
        X, yv = AdaProx.synth.dense_regression(50000, 2000, 0)
        A = np.asfortranarray(np.hstack([X, np.ones((50000, 1))]))
        nA = float(np.linalg.norm(A))
        kw = dict(f=O.Zero(), g=O.NormL1(10.0), h=O.Translate(O.NormL1(), -yv), A=A, eta=nA, t=1.0, tol=0.0)
        O.adaptive_linesearch_primal_dual(np.zeros(2001), np.zeros(50000), maxit=2, **kw)
        t0 = time.perf_counter()
        x, yy, it = O.adaptive_linesearch_primal_dual(np.zeros(2001), np.zeros(50000), maxit=30, **kw)
        line("C3 lad 50000x2001 AdaPDM+ (30 iterations)", it, time.perf_counter() - t0)
    if "svm" in which:
        N, d = 20000, 2000
        X, y = AdaProx.synth.dense_classification(N, d, 0)
        Z = y[:, None] * X
        Q = np.asfortranarray(Z @ Z.T)
        A = y[None, :].copy()
        kw = dict(f=O.Quadratic(Q, -np.ones(N)), g=O.IndBox(0.0, 0.1), h=O.IndZero(), A=A, rule=O.OurRule(t=0.1, norm_A=float(np.sqrt(N))), tol=0.0)
        O.adaptive_primal_dual(np.zeros(N), np.zeros(1), maxit=2, **kw)
        t0 = time.perf_counter()
        x, yy, it = O.adaptive_primal_dual(np.zeros(N), np.zeros(1), maxit=30, **kw)
        line(f"C3 dual SVM N={N} dense Q AdaPDM t=0.1 (30 iterations)", it, time.perf_counter() - t0)


if __name__ == "__main__":
    main()
