"""Regenerates tests/golden/*.json from the CPU oracle (oracle/adaprox_oracle.py).

The reference ships no golden vectors and cannot run here (no Julia), so these
fixtures pin the ORACLE's behaviour on deterministic, RNG-free or
counter-RNG problems; the oracle itself is pinned against first principles in
tests/test_oracle_known_answers.py.  Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import adaprox_oracle as O          # noqa: E402
import adaprox_b200                              # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def fl(seq):
    return [float(v) for v in seq]


def main():
    out = {}
    # 1. Simple2D (test/runtests.jl:6-51)
    f, g = O.Simple2DObjective(), O.Simple2DBox()
    log = []
    sol, it = O.adaptive_proxgrad(np.ones(2), f=f, g=g, rule=O.OurRule(gamma=1.0), log=log)
    out["simple2d_adapgm"] = dict(it=it, sol=fl(sol), f=float(f(sol)), gamma=fl(r["gamma"] for r in log[:12]),
                                  norm_res=fl(r["norm_res"] for r in log[:12]))
    sol, it = O.backtracking_proxgrad(np.ones(2), f=f, g=g, gamma0=1.0, xi=1.1)
    out["simple2d_backtracking"] = dict(it=it, sol=fl(sol), f=float(f(sol)))
    sol, it = O.backtracking_nesterov(np.ones(2), f=f, g=g, gamma0=1.0)
    out["simple2d_nesterov"] = dict(it=it, sol=fl(sol), f=float(f(sol)))
    # 2. Nesterov worst case (nesterov_worst_case/runme.jl:42-56), first 3000 iterations
    fw = O.WorstQuadratic(100, 100.0)
    for nm, rule in (("our", O.OurRule(gamma=0.01)), ("mm", O.MalitskyMishchenkoRule(gamma=0.01)), ("fixed", O.FixedStepsize(0.01))):
        log = []
        sol, it = O.adaptive_proxgrad(np.zeros(100), f=fw, g=O.Zero(), rule=rule, tol=1e-6, maxit=3000, log=log)
        out["worst_" + nm] = dict(it=it, f=float(fw(sol)), gamma=fl(r["gamma"] for r in log[:12]),
                                  objective_at=fl(log[k]["objective"] for k in (0, 9, 99, 999, 2999)))
    # 3. planted lasso 400 x 1000 (config C1), counter-based RNG
    P = adaprox_b200.synth.planted_lasso(400, 1000, 5, 0)
    Lf = adaprox_b200.synth.spectral_norm_sq(P["A"], iters=1000, tol=1e-15)
    fo = O.Counting(O.LinearLeastSquares(P["A"], P["b"]))
    log = []
    sol, it = O.adaptive_proxgrad(np.zeros(1000), f=fo, g=O.NormL1(1.0), rule=O.OurRule(gamma=1 / Lf), tol=1e-6, maxit=10000, log=log)
    out["lasso_c1_our"] = dict(it=it, Lf=Lf, optimum=P["optimum"], objective=float(log[-1]["objective"]),
                               gamma=fl(r["gamma"] for r in log[:40]), norm_res=fl(r["norm_res"] for r in log[:40]),
                               objective_prefix=fl(r["objective"] for r in log[:40]),
                               A_checksum=float(np.sum(P["A"] * np.cos(np.arange(P["A"].size).reshape(P["A"].shape)))),
                               b_head=fl(P["b"][:4]), eval_count=fo.eval_count, grad_count=fo.grad_count)
    with open(os.path.join(HERE, "oracle_golden.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", os.path.join(HERE, "oracle_golden.json"))


if __name__ == "__main__":
    main()
