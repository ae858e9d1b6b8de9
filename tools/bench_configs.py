"""Timings of the other BASELINE.json configs on one GPU (C1 lasso 400x1000, C2 sparse logreg rcv1 shape,
C3 LAD / sqrt-lasso 50000x2001 with AdaPDM+, dual SVM with AdaPDM in the dense-Q and the Gram form).  Prints one JSON line per config.
These are parity-test cases, not the headline bench; numbers go to profiles/."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import adaprox_b200 as AdaProx  # noqa: E402


def timed(fn):
    t0 = time.perf_counter()
    out = fn()
    return out, time.perf_counter() - t0


def main():
    which = sys.argv[1:] or ["c1", "c2", "lad", "sqrtlasso", "svm", "c5"]
    AdaProx.default_device()
    if "c1" in which:
        P = AdaProx.synth.planted_lasso(400, 1000, 5, 0)
        Lf = AdaProx.synth.spectral_norm_sq(P["A"], iters=1000, tol=1e-15)
        f, g = AdaProx.LinearLeastSquares(P["A"], P["b"]), AdaProx.NormL1(1.0)
        for tol in (1e-5, 1e-6):
            AdaProx.adaptive_proxgrad(np.zeros(1000), f=f, g=g, rule=AdaProx.OurRule(gamma=1 / Lf), tol=tol, maxit=10000)
            (x, it), wall = timed(lambda: AdaProx.adaptive_proxgrad(np.zeros(1000), f=f, g=g, rule=AdaProx.OurRule(gamma=1 / Lf), tol=tol, maxit=10000))
            info = AdaProx.last_solve_info()
            print(json.dumps(dict(config="C1 lasso 400x1000 AdaPGM OurRule", tol=tol, iterations=it, device_ms=info["solve_ms"],
                                  us_per_iteration=1e3 * info["solve_ms"] / it, iters_per_s=it / (info["solve_ms"] * 1e-3),
                                  e2e_ms=wall * 1e3, objective_gap=float(f(x) + g(x) - P["optimum"]))), flush=True)
        (x, it), wall = timed(lambda: AdaProx.fixed_proxgrad(np.zeros(1000), f=f, g=g, gamma=1 / Lf, tol=1e-5, maxit=10000))
        info = AdaProx.last_solve_info()
        print(json.dumps(dict(config="C1 lasso 400x1000 fixed-step PGM", tol=1e-5, iterations=it, device_ms=info["solve_ms"],
                              us_per_iteration=1e3 * info["solve_ms"] / it)), flush=True)
    if "c2" in which:
        import scipy.sparse as sp
        (rp, ci, va, y), tg = timed(lambda: AdaProx.synth.sparse_logreg(20242, 47236, 0))
        X = sp.csr_matrix((va, ci, rp), shape=(20242, 47236))
        n = 47237
        gam = 4 * 20242 / (va @ va + 20242)            # 4m / ||[X 1]||_F^2  (SURVEY 8d substitute for the m x m Gram)
        f = AdaProx.LogisticLoss(X, y)
        g = AdaProx.NormL1(0.03 * AdaProx.synth.logreg_lambda_max(X, y))     # 0.03 lambda_max: where the reference's lam = 0.01 sits on its own data sets (runme.jl:182)
        AdaProx.adaptive_proxgrad(np.zeros(n), f=f, g=g, rule=AdaProx.OurRule(gamma=gam), tol=1e-7, maxit=50)
        for tol in (1e-6, 1e-7):
            (x, it), wall = timed(lambda: AdaProx.adaptive_proxgrad(np.zeros(n), f=f, g=g, rule=AdaProx.OurRule(gamma=gam), tol=tol, maxit=2000))
            info = AdaProx.last_solve_info()
            nnz = len(va)
            bytes_iter = 2 * 12 * nnz + 8 * (4 * 20242 + 8 * n)
            print(json.dumps(dict(config="C2 sparse logreg 20242x47236 CSR AdaPGM", nnz=nnz, tol=tol, iterations=it, device_ms=info["solve_ms"],
                                  us_per_iteration=1e3 * info["solve_ms"] / it, iters_per_s=it / (info["solve_ms"] * 1e-3),
                                  l2_gbs=bytes_iter * it / (info["solve_ms"] * 1e-3) / 1e9, final_norm_res=info["final_norm_res"],
                                  nnz_w=int(np.count_nonzero(x)), gen_s=tg)), flush=True)
    for name in ("lad", "sqrtlasso"):
        if name not in which:
            continue
        (X, yv), tg = timed(lambda: AdaProx.synth.dense_regression(50000, 2000, 0))
        A = np.hstack([X, np.ones((50000, 1))])
        nA = float(np.linalg.norm(A))
        Ad = AdaProx.Counting(AdaProx.DeviceMatrix(A))
        h = AdaProx.Translate(AdaProx.NormL1() if name == "lad" else AdaProx.NormL2(), -yv)
        kw = dict(f=AdaProx.Zero(), g=AdaProx.NormL1(10.0), h=h, A=Ad, eta=nA, t=1.0, tol=1e-5, maxit=5000 if name == "lad" else 2000)
        AdaProx.adaptive_linesearch_primal_dual(np.zeros(2001), np.zeros(50000), **{**kw, "maxit": 20})
        Ad.mul_count = Ad.amul_count = 0
        (res, wall) = timed(lambda: AdaProx.adaptive_linesearch_primal_dual(np.zeros(2001), np.zeros(50000), **kw))
        x, yy, it = res
        info = AdaProx.last_solve_info()
        passes = Ad.mul_count + Ad.amul_count
        print(json.dumps(dict(config=f"C3 {name} 50000x2001 AdaPDM+", iterations=it, device_ms=info["solve_ms"],
                              us_per_iteration=1e3 * info["solve_ms"] / it, matrix_passes=passes,
                              hbm_gbs=passes * 50000 * 2001 * 8 / (info["solve_ms"] * 1e-3) / 1e9,
                              final_norm_res=info["final_norm_res"], gen_s=tg)), flush=True)
    if "svm" in which:
        N, d = 20000, 2000
        (X, y), tg = timed(lambda: AdaProx.synth.dense_classification(N, d, 0))
        Q, tq = timed(lambda: (y[:, None] * X) @ (X.T * y[None, :]))
        f = AdaProx.Quadratic(Q, -np.ones(N))
        Amat = AdaProx.DeviceMatrix(y[None, :].copy())
        for t in (0.1, 1.0):
            (res, wall) = timed(lambda: AdaProx.adaptive_primal_dual(np.zeros(N), np.zeros(1), f=f, g=AdaProx.IndBox(0.0, 0.1), h=AdaProx.IndZero(),
                                                                   A=Amat, rule=AdaProx.OurRule(t=t, norm_A=float(np.sqrt(N))), tol=1e-5, maxit=10000))
            x, yy, it = res
            info = AdaProx.last_solve_info()
            print(json.dumps(dict(config=f"C3 dual SVM N={N} dense Q AdaPDM t={t}", iterations=it, device_ms=info["solve_ms"],
                                  us_per_iteration=1e3 * info["solve_ms"] / it, hbm_gbs=it * N * N * 8 / (info["solve_ms"] * 1e-3) / 1e9,
                                  final_norm_res=info["final_norm_res"], gen_s=tg + tq)), flush=True)
    if "svmgram" in which:
        # the same dual SVM with Q = Z Z' never formed (QuadraticGram, Z = Dy X): two sweeps over Z per iteration (16 N d bytes)
        # instead of one over Q (8 N^2).  N = 20000 repeats the dense-Q instance above; N = 50000 is BASELINE configs[2]'s size
        # (the dense Q would be 20 GB and ~4 minutes of host DGEMM to build).
        for N, d in ((20000, 2000), (50000, 2000)):
            (X, y), tg = timed(lambda: AdaProx.synth.dense_classification(N, d, 0))
            Zm = AdaProx.DeviceMatrix(y[:, None] * X)
            f = AdaProx.QuadraticGram(Zm, -np.ones(N))
            Amat = AdaProx.DeviceMatrix(y[None, :].copy())
            for t in (0.1, 1.0):
                (res, wall) = timed(lambda: AdaProx.adaptive_primal_dual(np.zeros(N), np.zeros(1), f=f, g=AdaProx.IndBox(0.0, 0.1), h=AdaProx.IndZero(),
                                                                       A=Amat, rule=AdaProx.OurRule(t=t, norm_A=float(np.sqrt(N))), tol=1e-5, maxit=10000))
                x, yy, it = res
                info = AdaProx.last_solve_info()
                print(json.dumps(dict(config=f"C3 dual SVM N={N} d={d} Gram form AdaPDM t={t}", iterations=it, device_ms=info["solve_ms"],
                                      us_per_iteration=1e3 * info["solve_ms"] / it, hbm_gbs=it * 2 * N * d * 8 / (info["solve_ms"] * 1e-3) / 1e9,
                                      dense_q_equivalent_gbs=it * N * N * 8 / (info["solve_ms"] * 1e-3) / 1e9,
                                      final_norm_res=info["final_norm_res"], objective=float(f(x)), e2e_ms=wall * 1e3, gen_s=tg)), flush=True)
            Zm.free()
    if "c5" in which:
        # BASELINE configs[4]: batched multi-lambda lasso path, 256 lambdas, A 16384 x 8192, FP64 DMMA contractions.
        # Multi-GPU: the lambdas are split over the ranks (no collective), see tools/bench_path_multi.py.
        m, n, Lc = 16384, 8192, int(os.environ.get("C5_LAMBDAS", "256"))
        P, tg = timed(lambda: AdaProx.generate_planted_lasso(m, n, pfactor=5, seed=0, power_iters=30))
        f = AdaProx.LinearLeastSquares(P["A"], P["b"])
        lam_max = float(np.max(np.abs(P["A"].T @ P["b"].download())))
        lambdas = lam_max * (1e-3) ** (np.arange(Lc) / max(Lc - 1, 1))
        flop_iter = 4.0 * m * n * Lc                                     # two contractions of 2 m n L
        ms_r = P["A"].time_path_gemm(Lc, 0, reps=5)
        ms_g = P["A"].time_path_gemm(Lc, 1, reps=5)
        for maxit in (20, 300):
            (res, wall) = timed(lambda: AdaProx.adaptive_proxgrad_path(None, f=f, lambdas=lambdas, rule=AdaProx.OurRule(gamma=1 / P["Lf"]),
                                                                       tol=1e-6, maxit=maxit))
            X, its, info = res
            evals = info["batched_evals"]
            print(json.dumps(dict(config=f"C5 lambda path {m}x{n} L={Lc} AdaPGM OurRule (DMMA m8n8k4)", maxit=maxit, batched_evals=evals,
                                  device_ms=info["solve_ms"], ms_per_iteration=info["solve_ms"] / evals,
                                  tflops=flop_iter * evals / (info["solve_ms"] * 1e-3) / 1e12,
                                  gemm_AX_ms=ms_r, gemm_AtR_ms=ms_g, gemm_AX_tflops=2.0 * m * n * Lc / (ms_r * 1e-3) / 1e12,
                                  gemm_AtR_tflops=2.0 * m * n * Lc / (ms_g * 1e-3) / 1e12,
                                  converged_columns=int((info["norm_res"] <= 1e-6).sum()), iters_min=int(its.min()), iters_max=int(its.max()),
                                  nnz_first_last=[int((X[:, 0] != 0).sum()), int((X[:, -1] != 0).sum())], e2e_ms=wall * 1e3, gen_s=tg)), flush=True)


if __name__ == "__main__":
    main()
