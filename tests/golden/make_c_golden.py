"""Regenerates tests/golden/c_restatement_golden.json from the plain-C restatement (oracle/adaprox_ref.c), i.e. from an
implementation that shares no code with the numpy oracle.  tests/test_oracle_known_answers.py checks the numpy oracle against
these stored values, so the independent pin does not depend on gcc being present.  Run:  python tests/golden/make_c_golden.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import adaprox_b200                              # noqa: E402
from oracle import c_ref as R                    # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def fl(v):
    return [float(x) for x in v]


def cases():
    """name -> (callable returning (it, hist), description).  Shared with the test through this module."""
    P = adaprox_b200.synth.planted_lasso(100, 300, 10, 0)
    Lf = adaprox_b200.synth.spectral_norm_sq(P["A"], iters=1000, tol=1e-15)
    X, y = adaprox_b200.synth.dense_classification(120, 8, 0)
    Z = y[:, None] * X
    Q, q = Z @ Z.T, -np.ones(120)
    A1 = y[None, :].copy()
    rng = np.random.default_rng(2)
    A2 = np.hstack([rng.standard_normal((80, 6)), np.ones((80, 1))])
    b2 = A2 @ rng.standard_normal(7) + rng.laplace(size=80)
    return dict(P=P, Lf=Lf, Q=Q, q=q, A1=A1, A2=A2, b2=b2)


def main():
    D = cases()
    P, Lf = D["P"], D["Lf"]
    out = {}
    for nm, rule in (("our", R.RULE_OUR), ("mm", R.RULE_MM), ("plus", R.RULE_OUR_PLUS)):
        x, _, it, h = R.adaptive_primal_dual(np.zeros(300), None, f_kind=R.F_LEAST_SQUARES, F=P["A"], fvec=P["b"], g=R.prox_desc(R.P_NORM_L1, 1.0),
                                             rule=rule, gamma=1 / Lf, tol=1e-7, maxit=3000, nhist=30)
        out["lasso_100x300_" + nm] = dict(it=it, gamma=fl(h["gamma"]), norm_res=fl(h["norm_res"]), objective=fl(h["objective"]))
    nA = float(np.linalg.norm(D["A1"]))
    x, y, it, h = R.adaptive_primal_dual(np.zeros(120), np.zeros(1), f_kind=R.F_QUADRATIC, F=D["Q"], fvec=D["q"], g=R.prox_desc(R.P_IND_BOX, lo=0.0, hi=0.1),
                                         h=R.prox_desc(R.P_IND_ZERO), A=D["A1"], rule=R.RULE_OUR, gamma=1 / (2 * 1.2 * nA), t=1.0, norm_A=nA,
                                         tol=1e-6, maxit=5000, nhist=30)
    out["dual_svm_120"] = dict(it=it, gamma=fl(h["gamma"]), sigma=fl(h["sigma"]), norm_res=fl(h["norm_res"]))
    nA2 = float(np.linalg.norm(D["A2"]))
    for hn, kind in (("l1", R.P_NORM_L1), ("l2", R.P_NORM_L2)):
        x, y, it, h, trials = R.adaptive_linesearch_primal_dual(np.zeros(7), np.zeros(80), f_kind=R.F_ZERO, g=R.prox_desc(R.P_NORM_L1, 0.5),
                                                                h=R.prox_desc(kind, 1.0, shift=-D["b2"]), A=D["A2"], eta=0.05 * nA2, t=1.0,
                                                                tol=1e-6, maxit=3000, nhist=30)
        out["adapdm_plus_" + hn] = dict(it=it, trials=trials, gamma=fl(h["gamma"]), norm_res=fl(h["norm_res"]))
    x, it, h, ev = R.proxgrad_family(R.BACKTRACKING_NESTEROV, np.zeros(300), f_kind=R.F_LEAST_SQUARES, F=P["A"], fvec=P["b"],
                                     g=R.prox_desc(R.P_NORM_L1, 1.0), gamma=5.0 / Lf, tol=1e-7, maxit=1500, nhist=30)
    out["backtracking_nesterov"] = dict(it=it, evals=list(ev), gamma=fl(h["gamma"]), objective=fl(h["objective"]))
    x, y, it, h = R.malitsky_pock(np.zeros(120), np.zeros(1), f_kind=R.F_QUADRATIC, F=D["Q"], fvec=D["q"], g=R.prox_desc(R.P_IND_BOX, lo=0.0, hi=0.1),
                                  h=R.prox_desc(R.P_IND_ZERO), A=D["A1"], sigma=1 / nA, t=0.5, tol=1e-6, maxit=600, nhist=30)
    out["malitsky_pock"] = dict(it=it, gamma=fl(h["gamma"]), sigma=fl(h["sigma"]), norm_res=fl(h["norm_res"]))
    with open(os.path.join(HERE, "c_restatement_golden.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", os.path.join(HERE, "c_restatement_golden.json"))


if __name__ == "__main__":
    main()
