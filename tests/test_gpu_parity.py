"""Parity of the CUDA path (through the C ABI) against the CPU oracle.

Tolerances (BASELINE.json north_star): iterates and objective 1e-10 relative,
stepsize sequences 1e-12 relative, identical oracle-call counts.  The AdaPGM
trajectory is chaotic w.r.t. summation order, so whole runs are
compared by (a) single-call operator parity, (b) the free-running prefix of the
trajectory, (c) final objective / iterate / iteration count.
"""
import os

import numpy as np
import pytest

from oracle import adaprox_oracle as O
from oracle import drift

pytestmark = pytest.mark.gpu

RTOL_OP = 1e-13


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


# ---------------------------------------------------------------- operators
@pytest.mark.parametrize("m,n,order", [(1, 5, "C"), (7, 13, "F"), (400, 1000, "F"), (50, 2001, "C"),
                                       (33, 4100, "F"), (1000, 37, "C"), (9, 6200, "C")])
def test_mul_amul_dense(AdaProx, m, n, order):
    rng = np.random.default_rng(m * 1000 + n)
    A = np.asarray(rng.standard_normal((m, n)), order=order)
    x, y = rng.standard_normal(n), rng.standard_normal(m)
    M = AdaProx.DeviceMatrix(A)
    assert rel(M @ x, A @ x) < RTOL_OP * np.sqrt(n)
    assert rel(M.T @ y, A.T @ y) < RTOL_OP * np.sqrt(m)
    M.free()


def test_mul_amul_csr(AdaProx):
    import scipy.sparse as sp
    S = sp.random(300, 500, density=0.03, random_state=1, format="csc")
    S = S + sp.csr_matrix(([1.0], ([299], [499])), shape=(300, 500))
    rng = np.random.default_rng(2)
    x, y = rng.standard_normal(500), rng.standard_normal(300)
    M = AdaProx.DeviceMatrix(S)
    assert rel(M @ x, S @ x) < 1e-13
    assert rel(M.T @ y, S.T @ y) < 1e-13


def _pair(AdaProx, name, rng):
    import scipy.sparse as sp
    if name == "ls":
        A, b = np.asfortranarray(rng.standard_normal((60, 90))), rng.standard_normal(60)
        return AdaProx.LinearLeastSquares(A, b), O.LinearLeastSquares(A, b), 90
    if name == "ls_wide":
        A, b = rng.standard_normal((37, 4500)), rng.standard_normal(37)
        return AdaProx.LinearLeastSquares(A, b), O.LinearLeastSquares(A, b), 4500
    if name == "logistic_csr":
        X = sp.random(200, 80, density=0.1, random_state=3, format="csr")
        y = (rng.random(200) < 0.5).astype(float)
        return AdaProx.LogisticLoss(X, y), O.LogisticLoss(X, y), 81
    if name == "logistic_dense":
        X = rng.standard_normal((120, 30))
        y = (rng.random(120) < 0.5).astype(float)
        return AdaProx.LogisticLoss(X, y), O.LogisticLoss(X, y), 31
    if name == "quadratic":
        B = rng.standard_normal((70, 70)); Q = B @ B.T; q = rng.standard_normal(70)
        return AdaProx.Quadratic(Q, q), O.Quadratic(Q, q), 70
    if name == "quadratic_gram":                 # Quadratic(Z Z', q) by its factor; the oracle is the reference's dense-Q form
        Z = rng.standard_normal((70, 9)); q = rng.standard_normal(70)
        return AdaProx.QuadraticGram(Z, q), O.Quadratic(Z @ Z.T, q), 70
    if name == "quadratic_gram_wide":            # d spans three 2048-column chunks, ragged
        Z = rng.standard_normal((50, 4500)) / 60; q = rng.standard_normal(50)
        return AdaProx.QuadraticGram(np.asfortranarray(Z), q), O.Quadratic(Z @ Z.T, q), 50
    if name == "cubic":
        B = rng.standard_normal((40, 40)); Q = B @ B.T / 40; q = rng.standard_normal(40)
        return AdaProx.Cubic(Q, q, 0.7), O.Cubic(Q, q, 0.7), 40
    if name == "worst":
        return AdaProx.WorstQuadratic(100, 100.0), O.WorstQuadratic(100, 100.0), 120
    if name == "simple2d":
        return AdaProx.Simple2DObjective(), O.Simple2DObjective(), 2
    raise KeyError(name)


@pytest.mark.parametrize("name", ["ls", "ls_wide", "logistic_csr", "logistic_dense", "quadratic", "quadratic_gram", "quadratic_gram_wide",
                                  "cubic", "worst", "simple2d"])
def test_eval_with_pullback(AdaProx, name):
    rng = np.random.default_rng(5)
    fd, fo, n = _pair(AdaProx, name, rng)
    for _ in range(2):
        x = rng.standard_normal(n)
        vd, pbd = AdaProx.eval_with_pullback(fd, x)
        vo, pbo = O.eval_with_pullback(fo, x)
        assert abs(vd - vo) <= 1e-12 * max(abs(vo), 1.0)
        assert rel(pbd(), pbo()) < 1e-12


def test_prox_operators(AdaProx):
    rng = np.random.default_rng(7)
    x = rng.standard_normal(257) * 2
    b = rng.standard_normal(257)
    cases = [
        (AdaProx.NormL1(0.7), O.NormL1(0.7)),
        (AdaProx.NormL2(1.3), O.NormL2(1.3)),
        (AdaProx.NormL2(100.0), O.NormL2(100.0)),
        (AdaProx.IndBox(-0.3, 0.9), O.IndBox(-0.3, 0.9)),
        (AdaProx.Zero(), O.Zero()),
        (AdaProx.IndZero(), O.IndZero()),
        (AdaProx.Translate(AdaProx.NormL1(), -b), O.Translate(O.NormL1(), -b)),
        (AdaProx.Translate(AdaProx.NormL2(), -b), O.Translate(O.NormL2(), -b)),
    ]
    for gd, go in cases:
        for gamma in (0.05, 1.0, 3.7):
            yd, vd = AdaProx.prox(gd, x, gamma)
            yo, vo = O.prox(go, x, gamma)
            assert np.max(np.abs(yd - yo)) <= 1e-14 * max(1.0, np.max(np.abs(yo))), type(go).__name__
            assert abs(vd - vo) <= 1e-12 * max(1.0, abs(vo))
            # conjugates (Moreau), as the primal-dual loops use them
            yd, _ = AdaProx.prox(AdaProx.convex_conjugate(gd), x, gamma)
            yo, _ = O.prox(O.convex_conjugate(go), x, gamma)
            assert np.max(np.abs(yd - yo)) <= 1e-14 * max(1.0, np.max(np.abs(yo))), "conj " + type(go).__name__


def test_stepsize_rules(AdaProx):
    rng = np.random.default_rng(11)
    rules = [
        (AdaProx.FixedStepsize(0.3, 2.0), O.FixedStepsize(0.3, 2.0)),
        (AdaProx.MalitskyMishchenkoRule(0.3, 1.5), O.MalitskyMishchenkoRule(0.3, 1.5)),
        (AdaProx.OurRule(gamma=0.3), O.OurRule(gamma=0.3)),
        (AdaProx.OurRule(t=0.5, norm_A=2.0, delta=0.01), O.OurRule(t=0.5, norm_A=2.0, delta=0.01)),
        (AdaProx.OurRulePlus(gamma=0.3, nu=1.2, xi=0.9, r=0.6), O.OurRulePlus(gamma=0.3, nu=1.2, xi=0.9, r=0.6)),
    ]
    for rd, ro in rules:
        (g_d, s_d), st_d = AdaProx.stepsize(rd)
        (g_o, s_o), st_o = O.stepsize(ro)
        assert g_d == g_o and s_d == s_o
        for trial in range(20):
            x1, x0 = rng.standard_normal(50), rng.standard_normal(50)
            g1 = 3.0 * x1 + 0.1 * rng.standard_normal(50)
            g0 = 3.0 * x0 + 0.1 * rng.standard_normal(50)
            if trial == 7:
                g1 = g0.copy()               # dgrad = 0 -> NaN -> 0, third candidate Inf
            if st_o is None:
                continue
            (g_o, s_o), st_o2 = O.stepsize(ro, st_o, x1, g1, x0, g0)
            dg, dx = g1 - g0, x1 - x0
            (g_d, s_d), st_d2 = AdaProx.stepsize(rd, st_d, np.dot(dg, dg), np.dot(dg, dx), np.dot(dx, dx))
            assert abs(g_d - g_o) <= 1e-13 * abs(g_o), (type(ro).__name__, trial)
            assert abs(s_d - s_o) <= 1e-13 * abs(s_o)
            st_o, st_d = st_o2, st_d2


# ---------------------------------------------------------------- AdaPGM on the planted lasso (config C1)
def _run_both(AdaProx, P, rule_d, rule_o, tol, maxit=10_000):
    n = P["A"].shape[1]
    fd = AdaProx.Counting(AdaProx.LinearLeastSquares(P["A"], P["b"]))
    fo = O.Counting(O.LinearLeastSquares(P["A"], P["b"]))
    gd, go = AdaProx.Counting(AdaProx.NormL1(P["lam"])), O.Counting(O.NormL1(P["lam"]))
    logd, logo = [], []
    xd, itd = AdaProx.adaptive_proxgrad(np.zeros(n), f=fd, g=gd, rule=rule_d, tol=tol, maxit=maxit, log=logd)
    xo, ito = O.adaptive_proxgrad(np.zeros(n), f=fo, g=go, rule=rule_o, tol=tol, maxit=maxit, log=logo)
    return (xd, itd, logd, fd, gd), (xo, ito, logo, fo, go)


@pytest.mark.parametrize("rule", ["our", "mm", "fixed", "plus"])
def test_adapgm_lasso_c1(AdaProx, lasso_small, rule):
    P = lasso_small
    g0 = 1.0 / P["Lf"]
    rd, ro = {
        "our": (AdaProx.OurRule(gamma=g0), O.OurRule(gamma=g0)),
        "mm": (AdaProx.MalitskyMishchenkoRule(gamma=g0), O.MalitskyMishchenkoRule(gamma=g0)),
        "fixed": (AdaProx.FixedStepsize(g0), O.FixedStepsize(g0)),
        "plus": (AdaProx.OurRulePlus(gamma=g0), O.OurRulePlus(gamma=g0)),
    }[rule]
    maxit = 10_000 if rule != "fixed" else 600
    (xd, itd, logd, fd, gd), (xo, ito, logo, fo, go) = _run_both(AdaProx, P, rd, ro, 1e-6, maxit)
    # (b) free-running prefix.  The trajectory is chaotic w.r.t. rounding: ANY change
    # of summation order drifts.  The intrinsic drift is measured with the oracle itself on the same problem with
    # the columns of A permuted (identical in exact arithmetic); the CUDA path must stay within 1e-12 where the
    # oracle does, and inside a small multiple of the oracle's own drift envelope afterwards.
    K = 40
    perm = np.random.default_rng(0).permutation(P["A"].shape[1])
    logp = []
    _, itp = O.adaptive_proxgrad(np.zeros(1000), f=O.LinearLeastSquares(np.asfortranarray(P["A"][:, perm]), P["b"]), g=O.NormL1(P["lam"]),
                                 rule=ro, tol=1e-6, maxit=maxit, log=logp)
    gam_d = np.array([r["gamma"] for r in logd[:K]]); gam_o = np.array([r["gamma"] for r in logo[:K]])
    gam_p = np.array([r["gamma"] for r in logp[:K]])
    env = np.maximum.accumulate(np.abs(gam_p / gam_o - 1))
    drift = np.abs(gam_d / gam_o - 1)
    assert np.all(drift <= np.maximum(1e-12, 20 * env)), (drift.max(), env.max())
    assert np.max(drift[:15]) < 1e-12
    for key in ("norm_res", "objective"):
        a = np.array([r[key] for r in logd[:K]]); b_ = np.array([r[key] for r in logo[:K]])
        assert np.max(np.abs(a / b_ - 1)) < 1e-10, key
    # (c) final result
    obj_d = logd[-1]["objective"]; obj_o = logo[-1]["objective"]
    assert abs(obj_d - obj_o) <= 1e-10 * abs(obj_o)
    # iteration counts agree to within the oracle's own sensitivity to summation order (chaotic tail)
    assert abs(itd - ito) <= max(2, 0.03 * ito, 2 * abs(itp - ito)), (itd, ito, itp)
    if rule != "fixed":
        assert logd[-1]["norm_res"] <= 1e-6
        assert abs(obj_d - P["optimum"]) <= 1e-9 * P["optimum"]
        assert np.linalg.norm(xd - P["x_star"]) < 1e-5
    # counter identities -- exact
    assert fd.eval_count == itd + 1 and fd.grad_count == itd + 1
    assert gd.prox_count == (itd if logd[-1]["norm_res"] <= 1e-6 else itd + 1)
    assert [r["f_evals"] for r in logd[:5]] == [r["f_evals"] for r in logo[:5]] == [2, 3, 4, 5, 6]
    assert [r["prox_g_evals"] for r in logd[:5]] == [r["prox_g_evals"] for r in logo[:5]]


def test_teacher_forced_steps(AdaProx, lasso_small):
    """Step-wise parity deep into the run, independent of the chaotic drift: take the oracle's state at
    iterations 1, 10, 100, 500, 1000, 1500 and redo ONE iteration on the device from exactly that state."""
    P = lasso_small
    g0 = 1.0 / P["Lf"]
    trace = []
    O.adaptive_proxgrad(np.zeros(1000), f=O.LinearLeastSquares(P["A"], P["b"]), g=O.NormL1(1.0), rule=O.OurRule(gamma=g0),
                        tol=1e-6, maxit=1600, trace=trace)
    fd = AdaProx.LinearLeastSquares(P["A"], P["b"])
    gd = AdaProx.NormL1(1.0)
    rule = AdaProx.OurRule(gamma=g0)
    for k in (1, 10, 100, 500, 1000, 1500):
        if k + 1 >= len(trace):
            break
        prev, cur, nxt = trace[k - 1], trace[k], trace[k + 1]          # trace[i] is iteration i+1
        # gradient at the oracle's iterate
        f_d, pb = AdaProx.eval_with_pullback(fd, cur["x"])
        grad_d = pb()
        assert abs(f_d - cur["f_x"]) <= 1e-13 * abs(cur["f_x"])
        assert rel(grad_d, cur["grad_x"]) < 1e-13
        # stepsize formula from the oracle's state (gamma_k, gamma_{k-1}).  Near convergence |dgrad| << |grad|, so the
        # 1e-13 gradient difference would be amplified by |grad|/|dgrad|: the rule is checked on the oracle's vectors.
        dg, dx = cur["grad_x"] - prev["grad_x"], cur["x"] - prev["x"]
        gam_prev2 = trace[k - 2]["gamma"] if k >= 2 else g0
        (gam_d, _), _ = AdaProx.stepsize(rule, (prev["gamma"], gam_prev2), np.dot(dg, dg), np.dot(dg, dx), np.dot(dx, dx))
        assert abs(gam_d - cur["gamma"]) <= 1e-12 * cur["gamma"], k
        # prox-gradient step with the oracle's stepsize
        x_d, _ = AdaProx.prox(gd, cur["x"] - cur["gamma"] * grad_d, cur["gamma"])
        assert np.max(np.abs(x_d - nxt["x"])) <= 1e-13 * max(1.0, np.max(np.abs(nxt["x"])))


def test_adapgm_no_logger_same_result(AdaProx, lasso_small):
    """Without a logger the objective is never computed (src/AdaProx.jl:350-352) -- same iterates."""
    P = lasso_small
    f = AdaProx.LinearLeastSquares(P["A"], P["b"]); g = AdaProx.NormL1(1.0)
    rule = AdaProx.OurRule(gamma=1.0 / P["Lf"])
    log = []
    x1, it1 = AdaProx.adaptive_proxgrad(np.zeros(1000), f=f, g=g, rule=rule, tol=1e-5, maxit=300, log=log)
    x2, it2 = AdaProx.adaptive_proxgrad(np.zeros(1000), f=f, g=g, rule=rule, tol=1e-5, maxit=300)
    assert it1 == it2 == 300 and np.array_equal(x1, x2)          # deterministic reductions: bit-identical reruns
    assert log[0]["f_evals"] is None                              # f not wrapped in Counting -> `nothing`


# ---------------------------------------------------------------- reference test-suite mirror (test/runtests.jl)
def test_simple_2d_problem(AdaProx):
    f, g = AdaProx.Simple2DObjective(), AdaProx.Simple2DBox()
    fo = O.Simple2DObjective()
    obj_tol = 1e-7
    log = []
    sol, numit = AdaProx.adaptive_proxgrad(np.ones(2), f=f, g=g, rule=AdaProx.OurRule(gamma=1.0), log=log)
    assert fo(sol) < obj_tol and g(sol) == 0
    assert numit == 4472                                          # oracle / SURVEY section 4 item 1
    want = [0.025710976884666, 0.026039406387680, 0.036942695125715, 0.057454177251814, 0.069409118354550]
    assert np.allclose([r["gamma"] for r in log[:5]], want, rtol=1e-12)
    sol, numit = AdaProx.backtracking_proxgrad(np.ones(2), f=f, g=g, gamma0=1.0, xi=1.1)
    assert fo(sol) < obj_tol and g(sol) == 0
    so, no_ = O.backtracking_proxgrad(np.ones(2), f=fo, g=O.Simple2DBox(), gamma0=1.0, xi=1.1)
    assert numit == no_ and rel(sol, so) < 1e-9
    sol, numit = AdaProx.backtracking_nesterov(np.ones(2), f=f, g=g, gamma0=1.0)
    assert fo(sol) < obj_tol and g(sol) == 0
    so, no_ = O.backtracking_nesterov(np.ones(2), f=fo, g=O.Simple2DBox(), gamma0=1.0)
    assert numit == no_ and rel(sol, so) < 1e-9


def test_counting(AdaProx):
    f = AdaProx.Counting(AdaProx.Simple2DObjective())
    g = AdaProx.Counting(AdaProx.Simple2DBox())
    A = AdaProx.Counting(AdaProx.DeviceMatrix(np.eye(2)))
    x = np.ones(2)
    _, pb = AdaProx.eval_with_pullback(f, x)
    AdaProx.prox(g, x)
    A @ x
    assert (f.eval_count, f.grad_count, g.prox_count, A.mul_count, A.amul_count) == (1, 0, 1, 1, 0)
    pb()
    assert f.grad_count == 1
    A.T @ x
    assert A.amul_count == 1
    with AdaProx.without_counting():
        _, pb = AdaProx.eval_with_pullback(f, x)
        pb()
        AdaProx.prox(g, x)
        A @ x
    assert (f.eval_count, f.grad_count, g.prox_count, A.mul_count, A.amul_count) == (1, 1, 1, 1, 1)


def test_nesterov_worst_case(AdaProx):
    k = n = 100
    Lc = 100.0
    f, fo = AdaProx.WorstQuadratic(k, Lc), O.WorstQuadratic(k, Lc)
    fstar = (Lc / 8) * (1 / (k + 1) - 1)
    # (K, final-value tolerance): after 3000 iterations the adaptive trajectories have decorrelated (chaotic stepsizes),
    # only the fixed-step run stays comparable digit for digit
    for K, ftol, mk_d, mk_o in [(35, 2e-3, lambda: AdaProx.OurRule(gamma=1 / Lc), lambda: O.OurRule(gamma=1 / Lc)),
                                (15, 2e-3, lambda: AdaProx.MalitskyMishchenkoRule(gamma=1 / Lc), lambda: O.MalitskyMishchenkoRule(gamma=1 / Lc)),
                                (35, 1e-10, lambda: AdaProx.FixedStepsize(1 / Lc), lambda: O.FixedStepsize(1 / Lc))]:
        logd, logo = [], []
        xd, itd = AdaProx.adaptive_proxgrad(np.zeros(n), f=f, g=AdaProx.Zero(), rule=mk_d(), tol=1e-6, maxit=3000, log=logd)
        xo, ito = O.adaptive_proxgrad(np.zeros(n), f=fo, g=O.Zero(), rule=mk_o(), tol=1e-6, maxit=3000, log=logo)
        assert itd == ito == 3000
        gd = np.array([r["gamma"] for r in logd[:K]]); go = np.array([r["gamma"] for r in logo[:K]])
        assert np.max(np.abs(gd / go - 1)) < 1e-12
        assert abs(fo(xd) - fo(xo)) < ftol and fstar < fo(xd) < fstar + 0.06
    xd, itd = AdaProx.fixed_nesterov(np.zeros(n), f=f, g=AdaProx.Zero(), gamma=1 / Lc, tol=1e-6, maxit=2000)
    xo, ito = O.fixed_nesterov(np.zeros(n), f=fo, g=O.Zero(), gamma=1 / Lc, tol=1e-6, maxit=2000)
    assert itd == ito and rel(xd, xo) < 1e-9


# ---------------------------------------------------------------- proximal-gradient baselines on the lasso
def test_pg_baselines_lasso(AdaProx):
    P = AdaProx.synth.planted_lasso(100, 300, 10, 0)
    Lf = AdaProx.synth.spectral_norm_sq(P["A"], iters=1000, tol=1e-15)
    g0 = 1.0 / Lf
    n = 300
    def objs():
        return (AdaProx.Counting(AdaProx.LinearLeastSquares(P["A"], P["b"])), AdaProx.NormL1(1.0),
                O.Counting(O.LinearLeastSquares(P["A"], P["b"])), O.NormL1(1.0))
    for xi in (1.0, 1.5, 2.0):
        fd, gd, fo, go = objs(); ld, lo = [], []
        xd, itd = AdaProx.backtracking_proxgrad(np.zeros(n), f=fd, g=gd, gamma0=g0, xi=xi, tol=1e-7, maxit=400, log=ld)
        xo, ito = O.backtracking_proxgrad(np.zeros(n), f=fo, g=go, gamma0=g0, xi=xi, tol=1e-7, maxit=400, log=lo)
        K = 30
        assert np.allclose([r["gamma"] for r in ld[:K]], [r["gamma"] for r in lo[:K]], rtol=1e-12)
        assert [r["f_evals"] for r in ld[:K]] == [r["f_evals"] for r in lo[:K]]
        assert [r["grad_f_evals"] for r in ld[:K]] == [r["grad_f_evals"] for r in lo[:K]]
        assert np.allclose([r["objective"] for r in ld[:K]], [r["objective"] for r in lo[:K]], rtol=1e-10)
        assert abs(itd - ito) <= max(2, 0.03 * ito)
    fd, gd, fo, go = objs(); ld, lo = [], []
    xd, itd = AdaProx.backtracking_nesterov(np.zeros(n), f=fd, g=gd, gamma0=g0, tol=1e-7, maxit=400, log=ld)
    xo, ito = O.backtracking_nesterov(np.zeros(n), f=fo, g=go, gamma0=g0, tol=1e-7, maxit=400, log=lo)
    assert np.allclose([r["objective"] for r in ld[:30]], [r["objective"] for r in lo[:30]], rtol=1e-10)
    assert (fd.eval_count, fd.grad_count) == (fo.eval_count, fo.grad_count) or abs(itd - ito) <= 3
    fd, gd, fo, go = objs(); ld, lo = [], []
    xd, itd = AdaProx.fixed_nesterov(np.zeros(n), f=fd, g=gd, gamma=g0, tol=1e-7, maxit=400, log=ld)
    xo, ito = O.fixed_nesterov(np.zeros(n), f=fo, g=go, gamma=g0, tol=1e-7, maxit=400, log=lo)
    assert np.allclose([r["objective"] for r in ld[:30]], [r["objective"] for r in lo[:30]], rtol=1e-10)
    assert np.allclose([r["norm_res"] for r in ld[:30]], [r["norm_res"] for r in lo[:30]], rtol=1e-9)
    assert fd.eval_count == fo.eval_count                          # the logged f(x) is not counted
    fd, gd, fo, go = objs(); ld, lo = [], []
    x0 = np.random.default_rng(3).standard_normal(n)
    xd, itd = AdaProx.agraal(np.zeros(n), f=fd, g=gd, x0=x0, gamma0=g0, tol=1e-7, maxit=400, log=ld)
    xo, ito = O.agraal(np.zeros(n), f=fo, g=go, x0=x0, gamma0=g0, tol=1e-7, maxit=400, log=lo)
    assert np.allclose([r["gamma"] for r in ld[:30]], [r["gamma"] for r in lo[:30]], rtol=1e-11)
    assert np.allclose([r["objective"] for r in ld[:30]], [r["objective"] for r in lo[:30]], rtol=1e-10)


# ---------------------------------------------------------------- primal-dual: dual SVM, LAD, square-root lasso
def _svm(AdaProx, N=300, d=20, seed=0):
    X, y = AdaProx.synth.dense_classification(N, d, seed)
    Q = (y[:, None] * X) @ (X.T * y[None, :])
    return Q, -np.ones(N), y


def test_adapdm_dual_svm(AdaProx):
    Q, q, y = _svm(AdaProx)
    N = Q.shape[0]
    Amat = y[None, :].copy()
    nA = np.linalg.norm(Amat)
    for t in (0.1, 1.0):
        fd, fo = AdaProx.Counting(AdaProx.Quadratic(Q, q)), O.Counting(O.Quadratic(Q, q))
        Ad, Ao = AdaProx.Counting(AdaProx.DeviceMatrix(Amat)), O.Counting(Amat)
        ld, lo = [], []
        xd, yd, itd = AdaProx.adaptive_primal_dual(np.zeros(N), np.zeros(1), f=fd, g=AdaProx.IndBox(0.0, 0.1), h=AdaProx.IndZero(),
                                                   A=Ad, rule=AdaProx.OurRule(t=t, norm_A=nA), tol=1e-5, maxit=5000, log=ld)
        xo, yo, ito = O.adaptive_primal_dual(np.zeros(N), np.zeros(1), f=fo, g=O.IndBox(0.0, 0.1), h=O.IndZero(),
                                             A=Ao, rule=O.OurRule(t=t, norm_A=nA), tol=1e-5, maxit=5000, log=lo)
        K = 40
        assert np.allclose([r["gamma"] for r in ld[:K]], [r["gamma"] for r in lo[:K]], rtol=1e-12)
        assert np.allclose([r["sigma"] for r in ld[:K]], [r["sigma"] for r in lo[:K]], rtol=1e-12)
        assert np.allclose([r["norm_res"] for r in ld[:K]], [r["norm_res"] for r in lo[:K]], rtol=1e-9)
        assert abs(itd - ito) <= max(3, 0.05 * ito)
        assert np.all(xd >= 0) and np.all(xd <= 0.1) and abs(y @ xd) < 1e-4
        assert abs(fo.f(xd) - fo.f(xo)) <= 1e-7 * abs(fo.f(xo))
        assert (fd.eval_count, fd.grad_count, Ad.mul_count, Ad.amul_count) == (itd + 1, itd + 1, itd + 1, itd if ld[-1]["norm_res"] <= 1e-5 else itd + 1)
        assert [r["A_evals"] for r in ld[:5]] == [r["A_evals"] for r in lo[:5]]
        assert [r["At_evals"] for r in ld[:5]] == [r["At_evals"] for r in lo[:5]]


def test_dual_svm_gram_form(AdaProx):
    """dual_svm/runme.jl:47-59 with Q = Z Z' (Z = Dy X) never formed: `QuadraticGram(Z, q)` evaluates Q x as Z (Z' x) in all
    four persistent kernels.  The checker is the oracle on the reference's dense Q; on the CPU the two forms drift apart by
    < 1e-13 relative in the first 40 stepsizes (rounding only), identical iteration counts."""
    X, y = AdaProx.synth.dense_classification(300, 20, 0)
    Z = y[:, None] * X
    Q, q, N = Z @ Z.T, -np.ones(300), 300
    Amat = y[None, :].copy()
    nA = np.linalg.norm(Amat)
    fo0 = O.Quadratic(Q, q)
    for t in (0.1, 1.0):                                            # AdaPDM (k_primal_dual)
        fd, fo = AdaProx.Counting(AdaProx.QuadraticGram(Z, q)), O.Counting(O.Quadratic(Q, q))
        ld, lo = [], []
        xd, yd, itd = AdaProx.adaptive_primal_dual(np.zeros(N), np.zeros(1), f=fd, g=AdaProx.IndBox(0.0, 0.1), h=AdaProx.IndZero(),
                                                   A=AdaProx.DeviceMatrix(Amat), rule=AdaProx.OurRule(t=t, norm_A=nA), tol=1e-5, maxit=5000, log=ld)
        xo, yo, ito = O.adaptive_primal_dual(np.zeros(N), np.zeros(1), f=fo, g=O.IndBox(0.0, 0.1), h=O.IndZero(),
                                             A=Amat, rule=O.OurRule(t=t, norm_A=nA), tol=1e-5, maxit=5000, log=lo)
        K = 40
        assert np.allclose([r["gamma"] for r in ld[:K]], [r["gamma"] for r in lo[:K]], rtol=1e-11)
        assert np.allclose([r["norm_res"] for r in ld[:K]], [r["norm_res"] for r in lo[:K]], rtol=1e-8)
        assert np.allclose([r["objective"] for r in ld[:K]], [r["objective"] for r in lo[:K]], rtol=1e-10)
        assert abs(itd - ito) <= max(3, 0.05 * ito)
        assert abs(fo0(xd) - fo0(xo)) <= 1e-7 * abs(fo0(xo))
        assert (fd.eval_count, fd.grad_count) == (fo.eval_count + (itd - ito), fo.grad_count + (itd - ito))
    # AdaPGM, backtracking PG (k_proxgrad_family) and Malitsky-Pock (k_malitsky_pock) on the same term
    ld, lo = [], []
    xd, itd = AdaProx.adaptive_proxgrad(np.zeros(N), f=AdaProx.QuadraticGram(Z, q), g=AdaProx.IndBox(0.0, 0.1), rule=AdaProx.OurRule(gamma=1e-2), tol=1e-7, maxit=2000, log=ld)
    xo, ito = O.adaptive_proxgrad(np.zeros(N), f=O.Quadratic(Q, q), g=O.IndBox(0.0, 0.1), rule=O.OurRule(gamma=1e-2), tol=1e-7, maxit=2000, log=lo)
    assert np.allclose([r["gamma"] for r in ld[:40]], [r["gamma"] for r in lo[:40]], rtol=1e-11)
    assert abs(itd - ito) <= max(3, 0.05 * ito) and abs(fo0(xd) - fo0(xo)) <= 1e-9 * abs(fo0(xo))
    ld, lo = [], []
    xd, itd = AdaProx.backtracking_proxgrad(np.zeros(N), f=AdaProx.QuadraticGram(Z, q), g=AdaProx.IndBox(0.0, 0.1), gamma0=1.0, tol=1e-7, maxit=300, log=ld)
    xo, ito = O.backtracking_proxgrad(np.zeros(N), f=O.Quadratic(Q, q), g=O.IndBox(0.0, 0.1), gamma0=1.0, tol=1e-7, maxit=300, log=lo)
    assert np.allclose([r["gamma"] for r in ld[:30]], [r["gamma"] for r in lo[:30]], rtol=1e-12)
    assert np.allclose([r["objective"] for r in ld[:30]], [r["objective"] for r in lo[:30]], rtol=1e-10)
    ld, lo = [], []
    xd, yd, itd = AdaProx.malitsky_pock(np.zeros(N), np.zeros(1), f=AdaProx.QuadraticGram(Z, q), g=AdaProx.IndBox(0.0, 0.1), h=AdaProx.IndZero(),
                                        A=AdaProx.DeviceMatrix(Amat), t=1.0, sigma=1.0 / nA, tol=1e-5, maxit=300, log=ld)
    xo, yo, ito = O.malitsky_pock(np.zeros(N), np.zeros(1), f=O.Quadratic(Q, q), g=O.IndBox(0.0, 0.1), h=O.IndZero(),
                                  A=Amat, t=1.0, sigma=1.0 / nA, tol=1e-5, maxit=300, log=lo)
    assert np.allclose([r["gamma"] for r in ld[:30]], [r["gamma"] for r in lo[:30]], rtol=1e-11)
    assert np.allclose([r["norm_res"] for r in ld[:30]], [r["norm_res"] for r in lo[:30]], rtol=1e-8)
    # error behaviour: Z must have n rows
    with pytest.raises(AdaProx.AdaproxError):
        AdaProx.adaptive_proxgrad(np.zeros(N), f=AdaProx.QuadraticGram(Z[:-1], q), g=AdaProx.IndBox(0.0, 0.1), rule=AdaProx.OurRule(gamma=1e-2), maxit=3)


def test_dual_svm_gram_form_full_size(AdaProx):
    """BASELINE configs[2] size (N = 50000, d = 2000): the Gram-form value and gradient against numpy on the same factor, the
    affine structure of the gradient, and a short AdaPDM run that stays feasible and decreases the residual."""
    N, d = 50000, 2000
    X, y = AdaProx.synth.dense_classification(N, d, 0)
    Z = y[:, None] * X
    q = -np.ones(N)
    f = AdaProx.QuadraticGram(AdaProx.DeviceMatrix(Z), q)
    rng = np.random.default_rng(11)
    x = rng.random(N) * 0.1
    fx, pb = AdaProx.eval_with_pullback(f, x)
    gx = pb()
    t = Z @ (Z.T @ x)
    assert rel(gx, t + q) < 1e-11
    assert abs(fx - (0.5 * (x @ t) + x @ q)) <= 1e-11 * abs(0.5 * (x @ t) + x @ q)
    g2 = AdaProx.eval_with_pullback(f, 2.0 * x)[1]()
    assert rel(g2 - q, 2.0 * (gx - q)) < 1e-12                      # x -> Z Z' x is linear
    log = []
    xs, ys, it = AdaProx.adaptive_primal_dual(np.zeros(N), np.zeros(1), f=f, g=AdaProx.IndBox(0.0, 0.1), h=AdaProx.IndZero(),
                                              A=AdaProx.DeviceMatrix(y[None, :].copy()), rule=AdaProx.OurRule(t=0.1, norm_A=float(np.sqrt(N))),
                                              tol=1e-5, maxit=300, log=log)
    assert np.all(xs >= 0) and np.all(xs <= 0.1) and np.all(np.isfinite([r["gamma"] for r in log]))
    assert log[-1]["norm_res"] < 0.1 * log[0]["norm_res"]
    assert AdaProx.last_solve_info()["matrix_passes"] == 2


def test_sweep_direction_and_eviction_hints_leave_the_same_bits(AdaProx):
    """gemv.cuh / gemv_ring.cuh: A*x runs against the direction of the previous sweep over the same matrix and the ring's bulk
    copies carry an L2 evict_first policy.  Neither may change a bit: the Gram-form dual SVM (Z'x forward, Z*u reversed, one chunk)
    and LAD with AdaPDM+ on a 3-chunk matrix (A*x alternating, A'y forward) under every switch."""
    rng = np.random.default_rng(5)
    N, d = 3000, 700
    X = rng.standard_normal((N, d)) / np.sqrt(d)
    ysv = np.where(rng.random(N) < 0.5, -1.0, 1.0)
    m, n = 1500, 4500
    Al = rng.standard_normal((m, n)) / np.sqrt(n)
    bl = Al @ np.where(rng.random(n) < 0.05, rng.standard_normal(n), 0.0) + rng.laplace(scale=0.1, size=m)

    def solve():
        out = []
        Zm = AdaProx.DeviceMatrix(ysv[:, None] * X)                 # the switches are read when a matrix is uploaded
        Am = AdaProx.DeviceMatrix(ysv[None, :].copy())
        log = []
        xs, ys, it = AdaProx.adaptive_primal_dual(np.zeros(N), np.zeros(1), f=AdaProx.QuadraticGram(Zm, -np.ones(N)), g=AdaProx.IndBox(0.0, 0.1),
                                                  h=AdaProx.IndZero(), A=Am, rule=AdaProx.OurRule(t=0.1, norm_A=float(np.sqrt(N))),
                                                  tol=0.0, maxit=80, log=log)
        out += [xs.tobytes(), ys.tobytes(), np.array([r["gamma"] for r in log]).tobytes(), np.array([r["norm_res"] for r in log]).tobytes()]
        Zm.free(); Am.free()
        Ad = AdaProx.DeviceMatrix(Al)
        log = []
        xs, ys, it = AdaProx.adaptive_linesearch_primal_dual(np.zeros(n), np.zeros(m), f=AdaProx.Zero(), g=AdaProx.NormL1(0.5),
                                                             h=AdaProx.Translate(AdaProx.NormL1(), -bl), A=Ad, eta=float(np.linalg.norm(Al)), t=1.0,
                                                             tol=0.0, maxit=60, log=log)
        out += [xs.tobytes(), ys.tobytes(), np.array([r["gamma"] for r in log]).tobytes(), np.array([r["norm_res"] for r in log]).tobytes()]
        Ad.free()
        return out

    runs = {}
    for name, env in (("shipped", {}), ("one_way", {"ADAPROX_SWEEP_ONE_WAY": "1"}), ("no_hints", {"ADAPROX_L2_KEEP_MB": "-1"}),
                      ("keep_tail", {"ADAPROX_L2_KEEP_MB": "1"})):
        os.environ.update(env)
        try:
            runs[name] = solve()
        finally:
            for k in env:
                os.environ.pop(k, None)
    for name in ("one_way", "no_hints", "keep_tail"):
        assert runs[name] == runs["shipped"], name


def test_condat_vu(AdaProx):
    Q, q, y = _svm(AdaProx, 120, 10, 1)
    N = Q.shape[0]
    Amat = y[None, :].copy()
    Lf, nA = np.linalg.norm(Q), np.linalg.norm(Amat)
    ld, lo = [], []
    xd, yd, itd = AdaProx.condat_vu(np.zeros(N), np.zeros(1), f=AdaProx.Quadratic(Q, q), g=AdaProx.IndBox(0.0, 1.0), h=AdaProx.IndZero(),
                                    A=AdaProx.DeviceMatrix(Amat), Lf=Lf, norm_A=nA, tol=1e-5, maxit=300, log=ld)
    xo, yo, ito = O.condat_vu(np.zeros(N), np.zeros(1), f=O.Quadratic(Q, q), g=O.IndBox(0.0, 1.0), h=O.IndZero(),
                              A=Amat, Lf=Lf, norm_A=nA, tol=1e-5, maxit=300, log=lo)
    assert itd == ito
    assert np.allclose([r["norm_res"] for r in ld], [r["norm_res"] for r in lo], rtol=1e-8)
    assert rel(xd, xo) < 1e-9 and rel(yd, yo) < 1e-9


@pytest.mark.parametrize("hname", ["l1", "l2"])
def test_adapdm_plus_lad_sqrt_lasso(AdaProx, hname):
    X, yv = AdaProx.synth.dense_regression(200, 10, 0)
    m = X.shape[0]
    Amat = np.hstack([X, np.ones((m, 1))])
    nA = np.linalg.norm(Amat)
    lam = 0.1
    hd = AdaProx.Translate(AdaProx.NormL1() if hname == "l1" else AdaProx.NormL2(), -yv)
    ho = O.Translate(O.NormL1() if hname == "l1" else O.NormL2(), -yv)
    Ad, Ao = AdaProx.Counting(AdaProx.DeviceMatrix(Amat)), O.Counting(Amat)
    hdc, hoc = AdaProx.Counting(hd), O.Counting(ho)
    ld, lo, trials = [], [], []
    kw = dict(eta=nA, t=1.0, tol=1e-5, maxit=400)
    xd, yd, itd = AdaProx.adaptive_linesearch_primal_dual(np.zeros(11), np.zeros(m), f=AdaProx.Zero(), g=AdaProx.NormL1(lam), h=hdc, A=Ad, log=ld, **kw)
    xo, yo, ito = O.adaptive_linesearch_primal_dual(np.zeros(11), np.zeros(m), f=O.Zero(), g=O.NormL1(lam), h=hoc, A=Ao, log=lo, trials=trials, **kw)
    K = 30
    assert np.allclose([r["gamma"] for r in ld[:K]], [r["gamma"] for r in lo[:K]], rtol=1e-12)
    assert np.allclose([r["norm_res"] for r in ld[:K]], [r["norm_res"] for r in lo[:K]], rtol=1e-9)
    assert np.allclose([r["objective"] for r in ld[:K]], [r["objective"] for r in lo[:K]], rtol=1e-10)
    assert [r["At_evals"] for r in ld[:K]] == [r["At_evals"] for r in lo[:K]]        # same linesearch trial counts
    assert [r["prox_h_evals"] for r in ld[:K]] == [r["prox_h_evals"] for r in lo[:K]]
    assert abs(itd - ito) <= max(3, 0.05 * ito)
    # the generic loop with the same h
    ld, lo = [], []
    xd, yd, itd = AdaProx.adaptive_primal_dual(np.zeros(11), np.zeros(m), f=AdaProx.Zero(), g=AdaProx.NormL1(lam), h=hd, A=AdaProx.DeviceMatrix(Amat),
                                               rule=AdaProx.OurRule(t=1.0, norm_A=nA), tol=1e-5, maxit=200, log=ld)
    xo, yo, ito = O.adaptive_primal_dual(np.zeros(11), np.zeros(m), f=O.Zero(), g=O.NormL1(lam), h=ho, A=Amat,
                                         rule=O.OurRule(t=1.0, norm_A=nA), tol=1e-5, maxit=200, log=lo)
    assert np.allclose([r["norm_res"] for r in ld[:K]], [r["norm_res"] for r in lo[:K]], rtol=1e-9)
    assert np.allclose([r["objective"] for r in ld[:K]], [r["objective"] for r in lo[:K]], rtol=1e-10)


# ---------------------------------------------------------------- sparse logistic regression (config C2 shape, scaled)
def test_adapgm_sparse_logreg(AdaProx):
    import scipy.sparse as sp
    rp, ci, va, y = AdaProx.synth.sparse_logreg(m=600, n=900, seed=0, nnz_lo=10, nnz_hi=30)
    X = sp.csr_matrix((va, ci, rp), shape=(600, 900))
    n = 901
    X1 = sp.hstack([X, np.ones((600, 1))]).toarray()
    Lf = np.linalg.norm(X1 @ X1.T) / 4 / 600                       # sparse_logreg/runme.jl:58-59 (Frobenius)
    ld, lo = [], []
    xd, itd = AdaProx.adaptive_proxgrad(np.zeros(n), f=AdaProx.Counting(AdaProx.LogisticLoss(X, y)), g=AdaProx.NormL1(0.01),
                                        rule=AdaProx.OurRule(gamma=1 / Lf), tol=1e-7, maxit=2000, log=ld)
    xo, ito = O.adaptive_proxgrad(np.zeros(n), f=O.Counting(O.LogisticLoss(X, y)), g=O.NormL1(0.01),
                                  rule=O.OurRule(gamma=1 / Lf), tol=1e-7, maxit=2000, log=lo)
    K = 40
    assert np.allclose([r["gamma"] for r in ld[:K]], [r["gamma"] for r in lo[:K]], rtol=1e-12)
    assert np.allclose([r["objective"] for r in ld[:K]], [r["objective"] for r in lo[:K]], rtol=1e-10)
    assert abs(itd - ito) <= max(3, 0.05 * ito)
    assert abs(ld[-1]["objective"] - lo[-1]["objective"]) <= 1e-9 * abs(lo[-1]["objective"])


# ---------------------------------------------------------------- device generator vs host generator
def test_device_generator_matches_host(AdaProx):
    m, n = 64, 200
    Ph = AdaProx.synth.planted_lasso(m, n, 5, 3)
    Pd = AdaProx.generate_planted_lasso(m, n, 5, 3, power_iters=200)
    # raw uniforms are bit-identical; derived quantities differ only by summation order
    e = np.eye(n)
    cols = [0, 1, 57, n - 1]
    Ad = np.stack([Pd["A"] @ e[j] for j in cols], axis=1)
    assert rel(Ad, Ph["A"][:, cols]) < 1e-12
    assert rel(Pd["b"].download(), Ph["b"]) < 1e-12
    assert rel(Pd["x_star"], Ph["x_star"]) < 1e-12
    assert abs(Pd["optimum"] - Ph["optimum"]) < 1e-12 * Ph["optimum"]
    assert abs(Pd["Lf"] - np.linalg.norm(Ph["A"], 2) ** 2) < 1e-6 * Pd["Lf"]
    # planted solution is the minimiser: AdaPGM on the device-generated instance reaches the stated optimum
    f = AdaProx.LinearLeastSquares(Pd["A"], Pd["b"])
    log = []
    x, it = AdaProx.adaptive_proxgrad(np.zeros(n), f=f, g=AdaProx.NormL1(1.0), rule=AdaProx.OurRule(gamma=1 / Pd["Lf"]),
                                      tol=1e-8, maxit=20000, log=log)
    assert log[-1]["norm_res"] <= 1e-8
    assert abs(log[-1]["objective"] - Pd["optimum"]) < 1e-10 * Pd["optimum"]
    assert np.linalg.norm(x - Pd["x_star"]) < 1e-6


# ---------------------------------------------------------------- size-independent properties at a larger size
def test_large_gemv_properties(AdaProx):
    """Linearity and adjointness <A x, y> = <x, A'y> on a multi-chunk, multi-unit matrix (8192 x 12288)."""
    P = AdaProx.generate_planted_lasso(8192, 12288, 5, 1, power_iters=0)
    A = P["A"]
    rng = np.random.default_rng(0)
    x1, x2, y = rng.standard_normal(12288), rng.standard_normal(12288), rng.standard_normal(8192)
    Ax1, Ax2, Axs = A @ x1, A @ x2, A @ (2.0 * x1 - 3.0 * x2)
    assert rel(Axs, 2.0 * Ax1 - 3.0 * Ax2) < 1e-12
    assert abs(np.dot(Ax1, y) - np.dot(x1, A.T @ y)) <= 1e-11 * np.linalg.norm(Ax1) * np.linalg.norm(y)
    A.free()


def test_error_behaviour(AdaProx):
    with pytest.raises(ValueError):
        AdaProx.OurRule()                                         # src/AdaProx.jl:246
    with pytest.raises(ValueError):
        AdaProx.OurRulePlus()                                     # :288
    with pytest.raises(AssertionError):
        AdaProx.adaptive_linesearch_primal_dual(np.zeros(2), np.zeros(2), f=AdaProx.Zero(), g=AdaProx.Zero(), h=AdaProx.Zero(),
                                                A=np.eye(2), eta=-1.0)            # :481
    class Unknown:
        pass
    with pytest.raises(AdaProx.AdaproxError):
        AdaProx.adaptive_proxgrad(np.zeros(2), f=Unknown(), g=AdaProx.Zero(), rule=AdaProx.OurRule(gamma=1.0))
    with pytest.raises(AdaProx.AdaproxError):
        AdaProx.LinearLeastSquares(np.eye(3), np.ones(2))         # b shorter than the rows of A -> status < 0 at solve time
        AdaProx.adaptive_proxgrad(np.zeros(3), f=AdaProx.LinearLeastSquares(np.eye(3), np.ones(2)), g=AdaProx.Zero(),
                                  rule=AdaProx.OurRule(gamma=1.0))


# ---------------------------------------------------------------- Malitsky-Pock linesearch (src/AdaProx.jl:555-629)
def test_malitsky_pock(AdaProx):
    Q, q, y = _svm(AdaProx, 200, 12, 2)
    N = Q.shape[0]
    Amat = y[None, :].copy()
    nA = np.linalg.norm(Amat)
    for t in (0.5, 2.0):
        fd, fo = AdaProx.Counting(AdaProx.Quadratic(Q, q)), O.Counting(O.Quadratic(Q, q))
        Ad, Ao = AdaProx.Counting(AdaProx.DeviceMatrix(Amat)), O.Counting(Amat)
        gd, go = AdaProx.Counting(AdaProx.IndBox(0.0, 0.1)), O.Counting(O.IndBox(0.0, 0.1))
        ld, lo = [], []
        kw = dict(sigma=1 / nA, t=t, tol=1e-5, maxit=400)
        xd, yd, itd = AdaProx.malitsky_pock(np.zeros(N), np.zeros(1), f=fd, g=gd, h=AdaProx.IndZero(), A=Ad, log=ld, **kw)
        xo, yo, ito = O.malitsky_pock(np.zeros(N), np.zeros(1), f=fo, g=go, h=O.IndZero(), A=Ao, log=lo, **kw)
        K = 40
        assert np.allclose([r["sigma"] for r in ld[:K]], [r["sigma"] for r in lo[:K]], rtol=1e-12)
        assert np.allclose([r["gamma"] for r in ld[:K]], [r["gamma"] for r in lo[:K]], rtol=1e-12)
        assert np.allclose([r["norm_res"] for r in ld[:K]], [r["norm_res"] for r in lo[:K]], rtol=1e-8)
        for key in ("f_evals", "grad_f_evals", "prox_g_evals", "A_evals", "At_evals"):
            assert [r[key] for r in ld[:K]] == [r[key] for r in lo[:K]], key
        assert abs(itd - ito) <= max(3, 0.05 * ito)
    # LAD-type problem: h = |. - b|_1 through its conjugate, f = Zero
    X, yv = AdaProx.synth.dense_regression(150, 8, 1)
    A2 = np.hstack([X, np.ones((150, 1))])
    ld, lo = [], []
    kw = dict(sigma=1.0, t=1.0, tol=1e-5, maxit=150)
    AdaProx.malitsky_pock(np.zeros(9), np.zeros(150), f=AdaProx.Zero(), g=AdaProx.NormL1(0.1), h=AdaProx.Translate(AdaProx.NormL1(), -yv),
                          A=AdaProx.DeviceMatrix(A2), log=ld, **kw)
    O.malitsky_pock(np.zeros(9), np.zeros(150), f=O.Zero(), g=O.NormL1(0.1), h=O.Translate(O.NormL1(), -yv), A=A2, log=lo, **kw)
    assert np.allclose([r["norm_res"] for r in ld[:40]], [r["norm_res"] for r in lo[:40]], rtol=1e-8)
    assert np.allclose([r["objective"] for r in ld[:40]], [r["objective"] for r in lo[:40]], rtol=1e-10)


# ---------------------------------------------------------------- single-pass fused AdaPGM (solver_fused.cuh)
@pytest.mark.parametrize("m,n,pf", [(400, 1000, 5), (96, 9000, 60), (64, 20000, 200), (40, 131072, 2000)])
def test_fused_single_pass_matches_two_pass_and_oracle(AdaProx, m, n, pf):
    """Cluster sizes 1, 2 (ragged second CTA), 3 and 16; forced on with ADAPROX_FUSED=1 (small instances)."""
    import os
    P = AdaProx.synth.planted_lasso(m, n, pf, 4)
    Lf = AdaProx.synth.spectral_norm_sq(P["A"], iters=300)
    f_raw = AdaProx.LinearLeastSquares(P["A"], P["b"])
    runs = {}
    try:
        for mode in ("0", "1"):
            os.environ["ADAPROX_FUSED"] = mode
            os.environ["ADAPROX_RESIDENT"] = "0"          # mode 0 = the two-pass grid kernel (the 400 x 1000 case would run cluster-resident)
            os.environ["ADAPROX_GRIDRES"] = "0"           # ... or grid-resident
            f = AdaProx.Counting(f_raw)
            log = []
            x, it = AdaProx.adaptive_proxgrad(np.zeros(n), f=f, g=AdaProx.NormL1(1.0), rule=AdaProx.OurRule(gamma=1 / Lf),
                                              tol=1e-6, maxit=3000, log=log)
            runs[mode] = (x, it, log, AdaProx.last_solve_info())
    finally:
        os.environ.pop("ADAPROX_FUSED", None)
        os.environ.pop("ADAPROX_RESIDENT", None)
        os.environ.pop("ADAPROX_GRIDRES", None)
    logo = []
    xo, ito = O.adaptive_proxgrad(np.zeros(n), f=O.LinearLeastSquares(P["A"], P["b"]), g=O.NormL1(1.0), rule=O.OurRule(gamma=1 / Lf),
                                  tol=1e-6, maxit=3000, log=logo)
    # Free-running stepsize prefix, evidence-based: the distance of the device trajectory from the extended-precision run
    # of the same algorithm must stay inside 20 x the distance the FLOAT64 oracle itself (and two column permutations of
    # it: other valid summation orders) has from that run -- the intrinsic rounding drift, oracle/drift.py.  Where that
    # envelope is below 5e-14 the device is within 1e-12 of the extended-precision run, i.e. north_star's tolerance.
    K = 40
    truth = drift.lasso_runs(P["A"], P["b"], 1.0, lambda O_: O_.OurRule(gamma=1 / Lf), K, nperm=2, seed=m + n)
    env = drift.envelope(truth)
    for mode in ("0", "1"):
        x, it, log, info = runs[mode]
        ok, dd, allowed = drift.check_inside([r["gamma"] for r in log[:K]], truth, factor=20.0, floor=1e-12)
        assert ok, (mode, float(np.max(dd / allowed)), float(dd.max()), float(env.max()))
        kk = min(len(log), len(logo), 25)
        assert np.allclose([r["objective"] for r in log[:kk]], [r["objective"] for r in logo[:kk]], rtol=1e-10)
        # runs that hit maxit before converging are compared loosely (the tail of the trajectory is chaotic)
        ftol = 1e-10 if (it < 3000 and ito < 3000) else 1e-3
        assert abs(log[-1]["objective"] - logo[-1]["objective"]) <= ftol * abs(logo[-1]["objective"]), (mode, it, ito)
        assert abs(it - ito) <= max(3, 0.05 * ito)
        assert [r["f_evals"] for r in log[:5]] == [2, 3, 4, 5, 6]
    # one persistent launch for the whole solve (the helper CTAs only join sweeps of >= 4096 rows per cluster)
    assert runs["1"][3]["kernel_launches"] == 1


# ---------------------------------------------------------------- LIBSVM file -> CSR upload -> solve -> JSONL records
def test_libsvm_file_to_jsonl_records(AdaProx, tmp_path):
    import json
    import scipy.sparse as sp
    rp, ci, va, y = AdaProx.synth.sparse_logreg(m=300, n=500, seed=1, nnz_lo=5, nnz_hi=20)
    X = sp.csr_matrix((va, ci, rp), shape=(300, 500))
    path = str(tmp_path / "toy.libsvm")
    with open(path, "w") as fh:
        for i in range(300):
            feats = " ".join(f"{j + 1}:{float(v)!r}" for j, v in zip(X.indices[X.indptr[i]:X.indptr[i + 1]], X.data[X.indptr[i]:X.indptr[i + 1]]))
            fh.write(f"{'+1' if y[i] > 0.5 else '-1'} {feats}\n")
    Xl, yl = AdaProx.load_libsvm_dataset(path, labels=(0.0, 1.0))                 # sparse_logreg/runme.jl maps labels to {0, 1}
    assert Xl.shape[0] == 300 and np.array_equal(yl, y)
    Xl = sp.csr_matrix((Xl.data, Xl.indices, Xl.indptr), shape=(300, 500))        # trailing all-zero columns are not in the file
    assert abs(Xl - X).max() == 0.0
    n = 501
    X1 = sp.hstack([Xl, np.ones((300, 1))]).toarray()
    Lf = np.linalg.norm(X1 @ X1.T) / 4 / 300
    jl = str(tmp_path / "run.jsonl")
    lo = []
    with AdaProx.JsonlSink(jl, mode="w") as sink:
        xd, itd = AdaProx.adaptive_proxgrad(np.zeros(n), f=AdaProx.Counting(AdaProx.LogisticLoss(Xl, yl)), g=AdaProx.NormL1(0.01),
                                            rule=AdaProx.OurRule(gamma=1 / Lf), tol=1e-6, maxit=500, name="AdaPGM (1/Lf)", log=sink)
        AdaProx.fixed_proxgrad(np.zeros(n), f=AdaProx.Counting(AdaProx.LogisticLoss(Xl, yl)), g=AdaProx.NormL1(0.01), gamma=1 / Lf,
                               tol=1e-6, maxit=500, name="PGM (1/Lf)", log=sink)
    xo, ito = O.adaptive_proxgrad(np.zeros(n), f=O.Counting(O.LogisticLoss(Xl, yl)), g=O.NormL1(0.01), rule=O.OurRule(gamma=1 / Lf),
                                  tol=1e-6, maxit=500, log=lo)
    gb = AdaProx.read_jsonl(jl)
    assert list(gb) == ["AdaPGM (1/Lf)", "PGM (1/Lf)"]
    recs = gb["AdaPGM (1/Lf)"]
    assert len(recs) == itd and [r["it"] for r in recs] == list(range(1, itd + 1))
    assert list(recs[0]) == ["method", "it", "gamma", "sigma", "norm_res", "objective", "grad_f_evals", "prox_g_evals", "prox_h_evals",
                             "A_evals", "At_evals", "f_evals"]
    K = min(30, len(lo), len(recs))
    assert np.allclose([r["gamma"] for r in recs[:K]], [r["gamma"] for r in lo[:K]], rtol=1e-12)
    assert np.allclose([r["objective"] for r in recs[:K]], [r["objective"] for r in lo[:K]], rtol=1e-10)
    assert [r["grad_f_evals"] for r in recs[:5]] == [r["grad_f_evals"] for r in lo[:5]]
    # both reach the tolerance or not; the selection helper of logging.jl picks by gradient evaluations
    best = AdaProx.find_best(gb, list(gb), "norm_res", 1e-6, "grad_f_evals")
    assert best in gb
    with open(jl) as fh:
        json.loads(fh.readline())


# ---------------------------------------------------------------- batched multi-lambda lasso path (FP64 DMMA contractions)
@pytest.mark.parametrize("m,n,Lc", [(400, 1000, 7), (130, 300, 33), (257, 2050, 130)])
def test_lambda_path_matches_per_lambda_solves(AdaProx, m, n, Lc):
    """Column j of the batched solve == adaptive_proxgrad with NormL1(lambdas[j]) (the oracle runs them one by one).
    Shapes exercise ragged tiles in every dimension (rows, columns, lambdas not multiples of 128 / 16)."""
    P = AdaProx.synth.planted_lasso(m, n, 5, 2)
    Lf = AdaProx.synth.spectral_norm_sq(P["A"], iters=500)
    lam_max = float(np.max(np.abs(P["A"].T @ P["b"])))
    lambdas = lam_max * (1e-2) ** (np.arange(Lc) / max(Lc - 1, 1)) * 0.5
    f = AdaProx.LinearLeastSquares(P["A"], P["b"])
    H = 60
    X, its, info = AdaProx.adaptive_proxgrad_path(None, f=f, lambdas=lambdas, rule=AdaProx.OurRule(gamma=1 / Lf), tol=1e-6, maxit=400, history=H)
    Xo, itso, gh, rh, oh = O.adaptive_proxgrad_path(np.zeros((n, Lc)), f=O.LinearLeastSquares(P["A"], P["b"]), lambdas=lambdas,
                                                   rule_of=lambda j: O.OurRule(gamma=1 / Lf), tol=1e-6, maxit=400, history=H)
    K = 25
    gd, go = info["gamma_hist"][:K], gh[:K]
    live = ~np.isnan(go) & ~np.isnan(gd)
    assert np.array_equal(np.isnan(go[:12]), np.isnan(gd[:12]))                    # the same columns stop at the same early iterations
    # evidence-based drift bound per sampled column (first, middle, last lambda): device vs extended-precision run inside
    # 20 x the Float64 oracle's own drift envelope (oracle/drift.py), floor 1e-12
    for j in sorted({0, Lc // 2, Lc - 1}):
        kj = int(min(K, its[j], itso[j]))
        if kj < 3:
            continue
        truth = drift.lasso_runs(P["A"], P["b"], float(lambdas[j]), lambda O_: O_.OurRule(gamma=1 / Lf), kj, nperm=2, seed=j)
        ok, dd, allowed = drift.check_inside(info["gamma_hist"][:kj, j], truth, factor=20.0, floor=1e-12)
        assert ok, (j, float(np.max(dd / allowed)), float(dd.max()))
    od, oo = info["obj_hist"][:K], oh[:K]
    assert np.allclose(od[live], oo[live], rtol=1e-10)
    # stopping iterations agree within the oracle's own rounding sensitivity
    assert np.all(np.abs(its - itso) <= np.maximum(5, 0.1 * itso)), (its, itso)      # sanity for every column (20 of 354 seen on the widest case)
    # ... and evidence-based for the sampled columns: no further from the oracle than the oracle is from itself when the
    # columns of A are permuted (same algorithm, another Float64 summation order)
    perm = np.random.default_rng(1).permutation(n)
    Ap = np.asfortranarray(P["A"][:, perm])
    for j in sorted({0, Lc // 2, Lc - 1}):
        _, itp = O.adaptive_proxgrad(np.zeros(n), f=O.LinearLeastSquares(Ap, P["b"]), g=O.NormL1(float(lambdas[j])), rule=O.OurRule(gamma=1 / Lf),
                                     tol=1e-6, maxit=400)
        assert abs(int(its[j]) - int(itso[j])) <= max(3, 0.03 * itso[j], 2 * abs(itp - int(itso[j]))), (j, its[j], itso[j], itp)
    # columns that converged on both sides: the minimiser to O(tol), the objective to 1e-9; columns cut off at maxit are
    # compared through the objective only (their trajectories are chaotic w.r.t. rounding)
    conv = (its < 400) & (itso < 400)
    assert conv.sum() >= 1
    scale = np.maximum(np.linalg.norm(Xo, axis=0), 1e-12)
    assert np.max((np.linalg.norm(X - Xo, axis=0) / scale)[conv]) < 1e-4
    fo = np.array([0.5 * np.linalg.norm(P["A"] @ Xo[:, j] - P["b"]) ** 2 + lambdas[j] * np.abs(Xo[:, j]).sum() for j in range(Lc)])
    fd = np.array([0.5 * np.linalg.norm(P["A"] @ X[:, j] - P["b"]) ** 2 + lambdas[j] * np.abs(X[:, j]).sum() for j in range(Lc)])
    assert np.max((np.abs(fd - fo) / np.abs(fo))[conv]) < 1e-9
    assert np.max(np.abs(fd - fo) / np.abs(fo)) < 1e-2
    # larger lambda => sparser solution (a property of the path itself)
    nnz = (np.abs(X) > 0).sum(axis=0)
    assert nnz[0] <= nnz[-1]


def test_lambda_path_equals_single_solves_on_device(AdaProx):
    """Same device, same arithmetic family: the batched columns against adaptive_proxgrad called per lambda through the C ABI
    (two-pass GEMV kernel); also per-column gamma0 and a non-zero X0."""
    m, n, Lc = 200, 520, 9
    P = AdaProx.synth.planted_lasso(m, n, 5, 5)
    Lf = AdaProx.synth.spectral_norm_sq(P["A"], iters=500)
    rng = np.random.default_rng(0)
    lambdas = np.linspace(0.2, 2.0, Lc)
    gam0 = (1 / Lf) * np.linspace(0.5, 1.0, Lc)
    X0 = 0.01 * rng.standard_normal((n, Lc))
    f = AdaProx.LinearLeastSquares(P["A"], P["b"])
    X, its, info = AdaProx.adaptive_proxgrad_path(X0, f=f, lambdas=lambdas, rule=AdaProx.OurRule(gamma=1 / Lf), gamma0=gam0, tol=1e-5, maxit=6000)
    for j in range(Lc):
        xj, itj = AdaProx.adaptive_proxgrad(X0[:, j], f=f, g=AdaProx.NormL1(lambdas[j]), rule=AdaProx.OurRule(gamma=gam0[j]), tol=1e-5, maxit=6000)
        assert abs(int(its[j]) - itj) <= max(3, 0.05 * itj)
        assert np.linalg.norm(X[:, j] - xj) <= 5e-3 * max(np.linalg.norm(xj), 1e-9)          # both stopped at norm_res <= 1e-5
    assert np.all(info["norm_res"][its < 6000] <= 1e-5) and (its < 6000).sum() >= Lc // 2
    with pytest.raises(Exception):
        AdaProx.adaptive_proxgrad_path(None, f=f, lambdas=[-1.0], rule=AdaProx.OurRule(gamma=1 / Lf))


# ---------------------------------------------------------------- edge cases of the fused sweep and the lambda path
def test_fused_edge_cases(AdaProx):
    """Fewer rows than clusters, a single row, maxit = 1, chunk size 1 (every row its own chunk), and the fallback to the
    two-pass kernel when the row does not fit a 16-CTA cluster (n > 131072)."""
    import os
    rng = np.random.default_rng(3)

    def both(A, b, lam, gamma, maxit, env=None):
        res = {}
        for mode in ("0", "1"):
            os.environ["ADAPROX_FUSED"] = mode
            os.environ["ADAPROX_RESIDENT"] = "0"          # this test compares the sweep kernel with the two-pass grid kernel
            os.environ["ADAPROX_GRIDRES"] = "0"
            for k, v in (env or {}).items():
                os.environ[k] = v
            try:
                log = []
                x, it = AdaProx.adaptive_proxgrad(np.zeros(A.shape[1]), f=AdaProx.LinearLeastSquares(A, b), g=AdaProx.NormL1(lam),
                                                  rule=AdaProx.OurRule(gamma=gamma), tol=1e-9, maxit=maxit, log=log)
                res[mode] = (x, it, [r["gamma"] for r in log], [r["objective"] for r in log], AdaProx.last_solve_info()["matrix_passes"])
            finally:
                os.environ.pop("ADAPROX_FUSED", None)
                os.environ.pop("ADAPROX_RESIDENT", None)
                os.environ.pop("ADAPROX_GRIDRES", None)
                for k in (env or {}):
                    os.environ.pop(k, None)
        return res

    for (m, n, maxit, env) in [(3, 700, 40, None), (1, 64, 25, None), (50, 9000, 1, None), (37, 1200, 30, {"ADAPROX_FUSED_CHUNK": "1"}),
                               (64, 300, 30, {"ADAPROX_FUSED_CHUNK": "1000000"})]:
        A = rng.standard_normal((m, n)) / np.sqrt(n)
        b = rng.standard_normal(m)
        gam = 0.5 / np.linalg.norm(A, 2) ** 2
        R = both(A, b, 0.05, gam, maxit, env)
        assert R["0"][4] == 2 and R["1"][4] == 1
        assert R["0"][1] == R["1"][1]
        assert np.allclose(R["0"][2], R["1"][2], rtol=1e-10) and np.allclose(R["0"][3], R["1"][3], rtol=1e-11)
        assert np.allclose(R["0"][0], R["1"][0], rtol=1e-8, atol=1e-12)
        lo = []
        O.adaptive_proxgrad(np.zeros(n), f=O.LinearLeastSquares(A, b), g=O.NormL1(0.05), rule=O.OurRule(gamma=gam), tol=1e-9, maxit=maxit, log=lo)
        assert np.allclose(R["1"][2], [r["gamma"] for r in lo], rtol=1e-9)
    # n > 16 * 8192: not eligible, silently the two-pass kernel (same results)
    n = 131072 + 16
    A = np.zeros((2, n)); A[0, 0] = 1.0; A[1, n - 1] = 2.0; A[0, 5] = 0.5
    R = both(A, np.array([1.0, -1.0]), 0.01, 0.1, 20)
    assert R["1"][4] == 2 and np.array_equal(R["0"][0], R["1"][0])


def test_helper_ctas_leave_the_same_bits(AdaProx):
    """The helper CTAs beside the single-sweep kernel (solver_fused_helper.cuh: virtual clusters on the SMs no 16-CTA cluster can use, partial
    dots through global memory, rows streamed twice) take chunks from the same dispenser; whoever processes a chunk must leave the same bits.
    Full cluster width (n = 131072), tiny chunks and batches so that the helpers take part even in a short solve."""
    import os
    m, n = 192, 131072
    P = AdaProx.synth.planted_lasso(m, n, 2000, 6)
    Lf = AdaProx.synth.spectral_norm_sq(P["A"], iters=50)
    f = AdaProx.LinearLeastSquares(P["A"], P["b"])
    got = {}
    for tag, env in (("off", {"ADAPROX_HELPERS": "0"}), ("on", {"ADAPROX_HELPERS": "2", "ADAPROX_HELPER_HOLD": "0", "ADAPROX_HELPER_ROWS": "3"}),
                     ("on_b", {"ADAPROX_HELPERS": "1", "ADAPROX_HELPER_HOLD": "2", "ADAPROX_HELPER_ROWS": "16"})):
        os.environ.update(env, ADAPROX_FUSED="1", ADAPROX_FUSED_CHUNK="4")
        try:
            log = []
            x, it = AdaProx.adaptive_proxgrad(np.zeros(n), f=f, g=AdaProx.NormL1(1.0), rule=AdaProx.OurRule(gamma=1 / Lf), tol=0.0, maxit=300, log=log)
            got[tag] = (x, [r["gamma"] for r in log], [r["objective"] for r in log], AdaProx.last_solve_info())
        finally:
            for k in list(env) + ["ADAPROX_FUSED", "ADAPROX_FUSED_CHUNK"]:
                os.environ.pop(k, None)
    assert got["off"][3]["matrix_passes"] == 1 and got["off"][3]["kernel_launches"] == 1
    assert got["on"][3]["kernel_launches"] == 2                    # the helper launch went out
    for tag in ("on", "on_b"):
        assert np.array_equal(got["off"][0], got[tag][0]), tag
        assert got["off"][1] == got[tag][1] and got["off"][2] == got[tag][2], tag
    f.mat.free()


def test_lambda_path_edge_cases(AdaProx):
    """One lambda, one row, lambda = 0, sizes far from the tile sizes, maxit = 1; 65 lambdas (first width above the narrow tile)."""
    rng = np.random.default_rng(4)
    for (m, n, Lc, maxit) in [(1, 40, 1, 30), (33, 17, 3, 50), (70, 260, 65, 40), (129, 131, 2, 1)]:
        A = rng.standard_normal((m, n)) / np.sqrt(max(n, 1))
        b = rng.standard_normal(m)
        gam = 0.5 / max(np.linalg.norm(A, 2) ** 2, 1e-12)
        lambdas = np.linspace(0.0, 0.3, Lc)
        f = AdaProx.LinearLeastSquares(A, b)
        X, its, info = AdaProx.adaptive_proxgrad_path(None, f=f, lambdas=lambdas, rule=AdaProx.OurRule(gamma=gam), tol=1e-8, maxit=maxit, history=maxit)
        Xo, itso, gh, rh, oh = O.adaptive_proxgrad_path(np.zeros((n, Lc)), f=O.LinearLeastSquares(A, b), lambdas=lambdas,
                                                       rule_of=lambda j: O.OurRule(gamma=gam), tol=1e-8, maxit=maxit, history=maxit)
        K = min(maxit, 20)
        live = ~np.isnan(gh[:K]) & ~np.isnan(info["gamma_hist"][:K])
        assert live.any()
        assert np.allclose(info["gamma_hist"][:K][live], gh[:K][live], rtol=1e-9)
        assert np.allclose(info["obj_hist"][:K][live], oh[:K][live], rtol=1e-9, atol=1e-13)
        assert np.all(np.abs(its - itso) <= np.maximum(2, 0.1 * itso))
        ok = its == itso
        assert np.allclose(X[:, ok], Xo[:, ok], rtol=1e-6, atol=1e-9)


def test_fused_full_width_large_instance(AdaProx):
    """16384 x 131072 (17 GB, the full row width of BASELINE configs[3]: clusters of 16 CTAs, dynamic chunk schedule over
    7 clusters) generated on the device: the single-sweep kernel against the two-pass kernels on the same matrix, the planted
    optimum as the size-independent anchor, and run-to-run bit reproducibility of the dynamically scheduled sweep."""
    import os
    m, n = 16384, 131072
    dev = AdaProx.default_device()
    if dev.info()["free_bytes"] < 40e9:
        pytest.skip("needs ~20 GB of device memory")
    P = AdaProx.generate_planted_lasso(m, n, pfactor=5, seed=1, power_iters=20)
    f, g = AdaProx.LinearLeastSquares(P["A"], P["b"]), AdaProx.NormL1(1.0)
    runs = {}
    try:
        for mode in ("0", "1", "1"):
            os.environ["ADAPROX_FUSED"] = mode
            log = []
            x, it = AdaProx.adaptive_proxgrad(np.zeros(n), f=f, g=g, rule=AdaProx.OurRule(gamma=1 / P["Lf"]), tol=0.0, maxit=12, log=log)
            runs.setdefault(mode, []).append((x, log, AdaProx.last_solve_info()["matrix_passes"]))
    finally:
        os.environ.pop("ADAPROX_FUSED", None)
    (x2, l2, p2), (x1, l1, p1), (x1b, l1b, _) = runs["0"][0], runs["1"][0], runs["1"][1]
    assert p2 == 2 and p1 == 1
    assert np.allclose([r["gamma"] for r in l1], [r["gamma"] for r in l2], rtol=1e-9)
    assert np.allclose([r["objective"] for r in l1], [r["objective"] for r in l2], rtol=1e-11)
    assert np.allclose([r["norm_res"] for r in l1], [r["norm_res"] for r in l2], rtol=1e-9)
    assert np.linalg.norm(x1 - x2) <= 1e-9 * np.linalg.norm(x2)
    assert np.array_equal(x1, x1b) and [r["gamma"] for r in l1] == [r["gamma"] for r in l1b]      # dynamic schedule, same bits
    # the objective decreases towards the planted optimum and never undershoots it
    obj = np.array([r["objective"] for r in l1])
    assert np.all(obj >= P["optimum"] * (1 - 1e-12)) and obj[-1] < obj[0]
