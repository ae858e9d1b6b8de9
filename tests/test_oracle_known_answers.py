"""Pins the CPU oracle (oracle/adaprox_oracle.py).

The reference cannot run here (no Julia) and ships no golden vectors, so the
oracle is pinned against (1) the reference's own test inequalities
(test/runtests.jl), (2) first principles for every prox (brute-force argmin,
Moreau identity), (3) closed-form optima (Nesterov worst case, planted lasso
KKT), (4) the counter identities of src/counting.jl, and (5) the committed
fixtures tests/golden/oracle_golden.json (regenerate: tests/golden/make_golden.py).
"""
import json
import os

import numpy as np
import pytest
from scipy.optimize import minimize_scalar

from oracle import adaprox_oracle as O
import adaprox_b200

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.json")))


# ---------------------------------------------------------------- reference test file (test/runtests.jl)
def test_simple_2d_problem_reference_assertions():
    f, g = O.Simple2DObjective(), O.Simple2DBox()
    obj_tol = 1e-7
    sol, numit = O.adaptive_proxgrad(np.ones(2), f=f, g=g, rule=O.OurRule(gamma=1.0))
    assert f(sol) < obj_tol and g(sol) == 0                        # runtests.jl:35-36
    assert numit == GOLD["simple2d_adapgm"]["it"] == 4472
    sol, numit = O.backtracking_proxgrad(np.ones(2), f=f, g=g, gamma0=1.0, xi=1.1)
    assert f(sol) < obj_tol and g(sol) == 0                        # runtests.jl:42-43
    assert numit == GOLD["simple2d_backtracking"]["it"]
    sol, numit = O.backtracking_nesterov(np.ones(2), f=f, g=g, gamma0=1.0)
    assert f(sol) < obj_tol and g(sol) == 0                        # runtests.jl:49-50
    assert numit == GOLD["simple2d_nesterov"]["it"]


def test_counting_reference_assertions():
    f = O.Counting(O.Simple2DObjective())                          # runtests.jl:53-90
    g = O.Counting(O.Simple2DBox())
    A = O.Counting(np.eye(2))
    x = np.ones(2)
    _, pb = O.eval_with_pullback(f, x)
    O.prox(g, x)
    A @ x
    assert (f.eval_count, f.grad_count, g.prox_count, A.mul_count, A.amul_count) == (1, 0, 1, 1, 0)
    pb()
    assert f.grad_count == 1
    A.T @ x
    assert A.amul_count == 1
    with O.without_counting():
        _, pb = O.eval_with_pullback(f, x)
        pb()
        O.prox(g, x)
        A @ x
    assert (f.eval_count, f.grad_count, g.prox_count, A.mul_count, A.amul_count) == (1, 1, 1, 1, 1)


# ---------------------------------------------------------------- golden fixtures
def test_golden_simple2d_stepsizes():
    log = []
    O.adaptive_proxgrad(np.ones(2), f=O.Simple2DObjective(), g=O.Simple2DBox(), rule=O.OurRule(gamma=1.0), log=log)
    assert np.allclose([r["gamma"] for r in log[:12]], GOLD["simple2d_adapgm"]["gamma"], rtol=1e-14, atol=0)
    # SURVEY.md section 4 item 1 (independent probe of the survey session)
    assert np.allclose([r["gamma"] for r in log[:5]],
                       [0.025710976884666, 0.026039406387680, 0.036942695125715, 0.057454177251814, 0.069409118354550], rtol=1e-12)


@pytest.mark.parametrize("nm", ["our", "mm", "fixed"])
def test_golden_nesterov_worst_case(nm):
    fw = O.WorstQuadratic(100, 100.0)
    rule = {"our": O.OurRule(gamma=0.01), "mm": O.MalitskyMishchenkoRule(gamma=0.01), "fixed": O.FixedStepsize(0.01)}[nm]
    log = []
    sol, it = O.adaptive_proxgrad(np.zeros(100), f=fw, g=O.Zero(), rule=rule, tol=1e-6, maxit=3000, log=log)
    G = GOLD["worst_" + nm]
    assert it == G["it"]
    assert np.allclose([r["gamma"] for r in log[:12]], G["gamma"], rtol=1e-13)
    assert abs(fw(sol) - G["f"]) < 1e-10
    fstar = (100.0 / 8) * (1 / 101 - 1)                            # nesterov_worst_case/runme.jl:53
    assert fw(sol) > fstar and fw(sol) - fstar < 0.06
    if nm == "our":
        assert np.allclose([r["gamma"] for r in log[:4]], [0.014142135623731, 0.021973682269356, 0.035115112881319, 0.056600215833608], rtol=1e-12)


def test_golden_lasso_c1():
    G = GOLD["lasso_c1_our"]
    P = adaprox_b200.synth.planted_lasso(400, 1000, 5, 0)
    assert np.allclose(P["b"][:4], G["b_head"], rtol=1e-13)
    assert abs(np.sum(P["A"] * np.cos(np.arange(P["A"].size).reshape(P["A"].shape))) - G["A_checksum"]) < 1e-8
    fo = O.Counting(O.LinearLeastSquares(P["A"], P["b"]))
    log = []
    sol, it = O.adaptive_proxgrad(np.zeros(1000), f=fo, g=O.NormL1(1.0), rule=O.OurRule(gamma=1 / G["Lf"]), tol=1e-6, maxit=10000, log=log)
    assert np.allclose([r["gamma"] for r in log[:40]], G["gamma"], rtol=1e-11)
    assert abs(it - G["it"]) <= 0.03 * G["it"]
    assert abs(log[-1]["objective"] - G["objective"]) < 1e-10 * G["objective"]
    # planted optimum (lasso/runme.jl:77) and counter identities
    assert abs(log[-1]["objective"] - P["optimum"]) < 1e-9 * P["optimum"]
    assert np.linalg.norm(sol - P["x_star"]) < 1e-5
    assert fo.eval_count == it + 1 and fo.grad_count == it + 1


# ---------------------------------------------------------------- planted lasso: KKT by construction
def test_planted_lasso_kkt():
    P = adaprox_b200.synth.planted_lasso(60, 150, 5, 7)
    A, b, xs, ys, lam = P["A"], P["b"], P["x_star"], P["y_star"], P["lam"]
    r = A @ xs - b                                                  # = -y_star
    assert np.allclose(r, -ys, atol=1e-13)
    gradf = A.T @ r
    supp = xs != 0
    assert supp.sum() == int(150 / 5)
    assert np.allclose(gradf[supp], -lam * np.sign(xs[supp]), rtol=1e-10)      # -grad in lam * sign(x) on the support
    assert np.all(np.abs(gradf[~supp]) <= lam + 1e-12)                          # |grad| <= lam off the support
    assert abs(0.5 * np.dot(r, r) + lam * np.abs(xs).sum() - P["optimum"]) < 1e-12


# ---------------------------------------------------------------- prox operators from first principles
def _argmin_1d(fun, x, gamma):
    res = minimize_scalar(lambda y: fun(y) + (y - x) ** 2 / (2 * gamma), bounds=(-50, 50), method="bounded",
                          options={"xatol": 1e-12})
    return res.x


def test_prox_separable_vs_bruteforce():
    rng = np.random.default_rng(0)
    xs = rng.standard_normal(12) * 3
    for gamma in (0.1, 1.0, 2.5):
        y, v = O.prox(O.NormL1(0.7), xs, gamma)
        for xi, yi in zip(xs, y):
            assert abs(yi - _argmin_1d(lambda t: 0.7 * abs(t), xi, gamma)) < 1e-6
        assert abs(v - 0.7 * np.abs(y).sum()) < 1e-14
        y, v = O.prox(O.IndBox(-0.5, 1.25), xs, gamma)
        assert np.array_equal(y, np.clip(xs, -0.5, 1.25)) and v == 0
        y, v = O.prox(O.Zero(), xs, gamma)
        assert np.array_equal(y, xs) and v == 0
        y, v = O.prox(O.IndZero(), xs, gamma)
        assert not y.any() and v == 0


def test_prox_norml2_optimality():
    rng = np.random.default_rng(1)
    x = rng.standard_normal(9)
    for lam, gamma in ((1.0, 0.3), (2.0, 5.0)):
        y, v = O.prox(O.NormL2(lam), x, gamma)
        obj = lambda z: lam * np.linalg.norm(z) + np.dot(z - x, z - x) / (2 * gamma)
        for _ in range(200):                                        # no random perturbation improves the objective
            assert obj(y) <= obj(y + 1e-3 * rng.standard_normal(9)) + 1e-12
        assert abs(v - lam * np.linalg.norm(y)) < 1e-14
        if lam * gamma >= np.linalg.norm(x):
            assert not y.any()


def test_translate_and_moreau():
    rng = np.random.default_rng(2)
    w, b = rng.standard_normal(40) * 2, rng.standard_normal(40)
    for sigma in (0.2, 1.0, 4.0):
        # LAD: prox_{sigma h*} with h = |. - b|_1 is clamp(w - sigma b, -1, 1)        
        y, _ = O.prox(O.convex_conjugate(O.Translate(O.NormL1(), -b)), w, sigma)
        assert np.allclose(y, np.clip(w - sigma * b, -1, 1), atol=2e-15 * max(1, sigma))
        # sqrt-lasso: projection of w - sigma b onto the unit 2-norm ball
        y, _ = O.prox(O.convex_conjugate(O.Translate(O.NormL2(), -b)), w, sigma)
        z = w - sigma * b
        assert np.allclose(y, z / max(1.0, np.linalg.norm(z)), atol=1e-14 * max(1, sigma))
        # Moreau identity: prox_{sigma h*}(w) + sigma prox_{h / sigma}(w / sigma) = w
        for h in (O.NormL1(0.7), O.NormL2(1.3), O.IndBox(-0.3, 0.4), O.Translate(O.NormL1(2.0), b)):
            yc, _ = O.prox(O.convex_conjugate(h), w, sigma)
            yp, _ = O.prox(h, w / sigma, 1 / sigma)
            assert np.allclose(yc + sigma * yp, w, atol=1e-13)
    # the two direct specialisations
    y, _ = O.prox(O.convex_conjugate(O.Zero()), w, 0.7)
    assert not y.any()
    y, _ = O.prox(O.convex_conjugate(O.IndZero()), w, 0.7)
    assert np.array_equal(y, w)


# ---------------------------------------------------------------- Julia scalar semantics
def test_julia_scalar_semantics():
    assert np.isnan(O.jl_min(1.0, np.nan, 3.0)) and O.jl_min(2.0, 1.0, np.inf) == 1.0
    assert O.nan_to_zero(np.nan) == 0 and O.nan_to_zero(np.inf) == np.inf          # src/AdaProx.jl:24
    rule = O.OurRule(gamma=0.5)
    x1, x0 = np.array([1.0, 2.0]), np.array([0.0, 1.0])
    g = np.array([3.0, 3.0])
    # dgrad = 0: C = 0/0 -> NaN -> 0, L = 0, D = 0 -> third candidate gamma/0 = Inf; second is 1/0 = Inf
    (gam, sig), st = rule.stepsize((0.5, 0.5), x1, g, x0, g)
    assert gam == 0.5 * np.sqrt(2.0) and st == (gam, 0.5)
    with pytest.raises(ValueError):
        O.OurRule()
    with pytest.raises(ValueError):
        O.OurRulePlus()


# ---------------------------------------------------------------- primal-dual sanity on small instances
def test_adapdm_dual_svm_small():
    X, y = adaprox_b200.synth.dense_classification(300, 20, 0)
    Q = (y[:, None] * X) @ (X.T * y[None, :])
    A = y[None, :].copy()
    objs = []
    for t in (0.1, 1.0):
        f, Ac = O.Counting(O.Quadratic(Q, -np.ones(300))), O.Counting(A)
        log = []
        x, yy, it = O.adaptive_primal_dual(np.zeros(300), np.zeros(1), f=f, g=O.IndBox(0.0, 0.1), h=O.IndZero(), A=Ac,
                                           rule=O.OurRule(t=t, norm_A=np.linalg.norm(A)), tol=1e-5, maxit=10000, log=log)
        assert log[-1]["norm_res"] <= 1e-5 and it < 10000
        assert np.all(x >= 0) and np.all(x <= 0.1) and abs(y @ x) < 1e-4
        assert (f.eval_count, f.grad_count, Ac.mul_count, Ac.amul_count) == (it + 1, it + 1, it + 1, it)
        objs.append(f.f(x))
    assert abs(objs[0] - objs[1]) < 1e-5 * abs(objs[0])


@pytest.mark.parametrize("hname", ["l1", "l2"])
def test_adapdm_plus_small(hname):
    X, yv = adaprox_b200.synth.dense_regression(200, 10, 0)
    A = np.hstack([X, np.ones((200, 1))])
    h = O.Translate(O.NormL1() if hname == "l1" else O.NormL2(), -yv)
    trials, log = [], []
    x, y, it = O.adaptive_linesearch_primal_dual(np.zeros(11), np.zeros(200), f=O.Zero(), g=O.NormL1(0.1), h=O.Counting(h),
                                                 A=O.Counting(A), eta=np.linalg.norm(A), t=1.0, tol=1e-5, maxit=3000,
                                                 log=log, trials=trials)
    assert sum(trials) >= len(trials) and sum(trials) <= 1.5 * len(trials)      # ~1.07 trials per iteration
    obj = lambda z: 0.1 * np.abs(z).sum() + h(A @ z)
    if hname == "l2":
        assert log[-1]["norm_res"] <= 1e-5
        rng = np.random.default_rng(0)
        for _ in range(100):
            assert obj(x) <= obj(x + 1e-4 * rng.standard_normal(11)) + 1e-9
    assert log[-1]["objective"] <= log[0]["objective"]


def test_malitsky_pock_and_condat_vu_agree_with_adapdm():
    X, y = adaprox_b200.synth.dense_classification(120, 10, 1)
    Q = (y[:, None] * X) @ (X.T * y[None, :])
    A = y[None, :].copy()
    q = -np.ones(120)
    kw = dict(f=O.Quadratic(Q, q), g=O.IndBox(0.0, 1.0), h=O.IndZero(), A=A, tol=1e-6, maxit=20000)
    x1, _, it1 = O.adaptive_primal_dual(np.zeros(120), np.zeros(1), rule=O.OurRule(t=1.0, norm_A=np.linalg.norm(A)), **kw)
    x2, _, it2 = O.malitsky_pock(np.zeros(120), np.zeros(1), sigma=1 / np.linalg.norm(A), t=1.0, **kw)
    x3, _, it3 = O.condat_vu(np.zeros(120), np.zeros(1), Lf=np.linalg.norm(Q), norm_A=np.linalg.norm(A), **kw)
    fq = O.Quadratic(Q, q)
    assert abs(fq(x1) - fq(x2)) < 1e-5 * abs(fq(x1)) and abs(fq(x1) - fq(x3)) < 1e-5 * abs(fq(x1))
    assert max(it1, it2, it3) < 20000                              # all three converged within the budget


def test_baselines_reach_the_planted_optimum():
    P = adaprox_b200.synth.planted_lasso(100, 300, 10, 0)
    Lf = adaprox_b200.synth.spectral_norm_sq(P["A"], iters=1000, tol=1e-15)
    f, g = O.LinearLeastSquares(P["A"], P["b"]), O.NormL1(1.0)
    runs = {
        "bt": O.backtracking_proxgrad(np.zeros(300), f=f, g=g, gamma0=1 / Lf, xi=1.5, tol=1e-7, maxit=20000)[0],
        "nes": O.backtracking_nesterov(np.zeros(300), f=f, g=g, gamma0=1 / Lf, tol=1e-7, maxit=20000)[0],
        "fnes": O.fixed_nesterov(np.zeros(300), f=f, g=g, gamma=1 / Lf, tol=1e-7, maxit=20000)[0],
        "agraal": O.agraal(np.zeros(300), f=f, g=g, gamma0=1 / Lf, tol=1e-7, maxit=20000)[0],
        "auto": O.auto_adaptive_proxgrad(np.zeros(300), f=f, g=g, gamma=1 / Lf, tol=1e-7, maxit=20000)[0],
    }
    for nm, x in runs.items():
        assert abs(f(x) + g(x) - P["optimum"]) < 1e-8 * P["optimum"], nm


# ---------------------------------------------------------------- fixtures produced by the independent C restatement
def test_numpy_oracle_against_c_restatement_fixtures():
    """tests/golden/c_restatement_golden.json holds stepsize / residual / objective prefixes computed by oracle/adaprox_ref.c
    (generator: tests/golden/make_c_golden.py); the numpy oracle must reproduce them.  Needs no compiler."""
    import importlib.util
    import json as _json
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_c_golden", os.path.join(here, "make_c_golden.py"))
    try:
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)                                # imports oracle.c_ref (ctypes only; nothing is built or loaded)
    except Exception as e:                                          # pragma: no cover
        pytest.skip(f"cannot import the generator: {e}")
    G = _json.load(open(os.path.join(here, "c_restatement_golden.json")))
    D = mod.cases()
    P, Lf = D["P"], D["Lf"]

    def close(a, b, rtol):
        a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
        K = min(len(a), len(b))
        fin = np.isfinite(a[:K]) & np.isfinite(b[:K])
        return K >= 5 and np.array_equal(np.isfinite(a[:K]), np.isfinite(b[:K])) and np.allclose(a[:K][fin], b[:K][fin], rtol=rtol, atol=1e-13)

    for nm, rule in (("our", O.OurRule(gamma=1 / Lf)), ("mm", O.MalitskyMishchenkoRule(gamma=1 / Lf)), ("plus", O.OurRulePlus(gamma=1 / Lf))):
        log = []
        _, it = O.adaptive_proxgrad(np.zeros(300), f=O.LinearLeastSquares(P["A"], P["b"]), g=O.NormL1(1.0), rule=rule, tol=1e-7, maxit=3000, log=log)
        g = G["lasso_100x300_" + nm]
        assert close([r["gamma"] for r in log[:30]], g["gamma"], 1e-10) and close([r["norm_res"] for r in log[:30]], g["norm_res"], 1e-8)
        assert close([r["objective"] for r in log[:30]], g["objective"], 1e-10) and abs(it - g["it"]) <= max(3, 0.05 * g["it"])
    nA = float(np.linalg.norm(D["A1"]))
    log = []
    _, _, it = O.adaptive_primal_dual(np.zeros(120), np.zeros(1), f=O.Quadratic(D["Q"], D["q"]), g=O.IndBox(0.0, 0.1), h=O.IndZero(), A=D["A1"],
                                      rule=O.OurRule(t=1.0, norm_A=nA), tol=1e-6, maxit=5000, log=log)
    g = G["dual_svm_120"]
    assert close([r["gamma"] for r in log[:30]], g["gamma"], 1e-10) and close([r["sigma"] for r in log[:30]], g["sigma"], 1e-10)
    assert close([r["norm_res"] for r in log[:30]], g["norm_res"], 1e-8) and abs(it - g["it"]) <= max(3, 0.05 * g["it"])
    nA2 = float(np.linalg.norm(D["A2"]))
    for hn, hf in (("l1", O.NormL1()), ("l2", O.NormL2())):
        Ao = O.Counting(D["A2"])
        log = []
        _, _, it = O.adaptive_linesearch_primal_dual(np.zeros(7), np.zeros(80), f=O.Zero(), g=O.NormL1(0.5), h=O.Translate(hf, -D["b2"]), A=Ao,
                                                     eta=0.05 * nA2, t=1.0, tol=1e-6, maxit=3000, log=log)
        g = G["adapdm_plus_" + hn]
        assert close([r["gamma"] for r in log[:30]], g["gamma"], 1e-10) and close([r["norm_res"] for r in log[:30]], g["norm_res"], 1e-8)
        if it == g["it"]:
            assert Ao.amul_count == 1 + g["trials"]
    fo = O.Counting(O.LinearLeastSquares(P["A"], P["b"]))
    log = []
    _, it = O.backtracking_nesterov(np.zeros(300), f=fo, g=O.NormL1(1.0), gamma0=5.0 / Lf, tol=1e-7, maxit=1500, log=log)
    g = G["backtracking_nesterov"]
    assert close([r["gamma"] for r in log[:30]], g["gamma"], 1e-12) and close([r["objective"] for r in log[:30]], g["objective"], 1e-10)
    if it == g["it"]:
        assert [fo.eval_count, fo.grad_count] == g["evals"]
    log = []
    O.malitsky_pock(np.zeros(120), np.zeros(1), f=O.Quadratic(D["Q"], D["q"]), g=O.IndBox(0.0, 0.1), h=O.IndZero(), A=D["A1"], sigma=1 / nA, t=0.5,
                    tol=1e-6, maxit=600, log=log)
    g = G["malitsky_pock"]
    assert close([r["gamma"] for r in log[:30]], g["gamma"], 1e-10) and close([r["sigma"] for r in log[:30]], g["sigma"], 1e-10)
    assert close([r["norm_res"] for r in log[:30]], g["norm_res"], 1e-8)


def test_extended_precision_mode_and_drift_envelope():
    """oracle.precision(np.longdouble) runs the same restatement in x87 extended precision; oracle/drift.py turns the distance of the Float64
    oracle (natural and permuted summation orders) from that run into the intrinsic-drift envelope the GPU parity tests use."""
    import adaprox_b200 as AdaProx
    from oracle import drift
    P = AdaProx.synth.planted_lasso(60, 150, 5, 1)
    Lf = AdaProx.synth.spectral_norm_sq(P["A"], iters=500)
    runs = drift.lasso_runs(P["A"], P["b"], 1.0, lambda O_: O_.OurRule(gamma=1 / Lf), 60, nperm=2, seed=3)
    assert isinstance(runs["ext"][5]["gamma"], np.longdouble) and isinstance(runs["f64"][5]["gamma"], np.float64)
    assert O.F64 is np.float64                                     # the mode is restored on exit
    env = drift.envelope(runs)
    assert len(env) == 60 and np.all(np.diff(env) >= 0)            # a running maximum
    assert env[4] < 1e-14 and 1e-12 < env[-1] < 1e-2               # Float64 pins the first iterations, then the trajectories drift apart
    # a Float64 run is inside its own envelope by construction; a perturbed one (gamma0 changed in the 13th digit) is not for long
    ok, dd, allowed = drift.check_inside([r["gamma"] for r in runs["perms"][0]], runs, factor=1.0, floor=0.0)
    assert ok
    logp = []
    O.adaptive_proxgrad(np.zeros(150), f=O.LinearLeastSquares(P["A"], P["b"]), g=O.NormL1(1.0), rule=O.OurRule(gamma=(1 / Lf) * (1 + 1e-9)), tol=0.0,
                        maxit=60, log=logp)
    ok, dd, allowed = drift.check_inside([r["gamma"] for r in logp], runs, factor=20.0, floor=1e-12)
    assert not ok


def test_adapgm_nan_stepsize_reaches_norm_res_through_the_dual_residual():
    """src/AdaProx.jl:342-348 with A = 0, h = Zero (AdaPGM, :418-421): when the rule returns NaN (Malitsky-Mishchenko with dx = 0:
    L = 0 / 0), w = y + sigma * ((1 + rho) * A_x - rho * A_x_prev) is NaN although A_x = 0, so dual_res and norm_res are NaN, the
    stopping test never fires and the next v = x - gamma * grad makes x NaN -- even though the primal residual alone is exactly 0.
    The restatement has to behave the same way (the device kernels are tested against it on this instance)."""
    m, n = 149, 3
    rng = np.random.default_rng(m * 31 + n)
    A = np.asfortranarray(rng.standard_normal((m, n)) / np.sqrt(m))
    b = rng.standard_normal(m)
    Lf = float(np.linalg.norm(A, 2) ** 2)
    rng.standard_normal(n)
    x0 = 0.05 * rng.standard_normal(n)
    log = []
    x, it = O.adaptive_proxgrad(x0, f=O.LinearLeastSquares(A, b), g=O.IndBox(-0.2, 0.5), rule=O.MalitskyMishchenkoRule(gamma=1 / Lf),
                                tol=1e-9, maxit=12, log=log)
    assert it == 12 and np.all(np.isnan(x))
    assert np.isfinite(log[0]["norm_res"]) and np.isfinite(log[0]["gamma"])
    assert all(np.isnan(r["norm_res"]) and np.isnan(r["gamma"]) for r in log[1:])
    # the same instance with a rule that cannot return NaN stops at the stationary point the box holds it in
    x2, it2 = O.adaptive_proxgrad(x0, f=O.LinearLeastSquares(A, b), g=O.IndBox(-0.2, 0.5), rule=O.FixedStepsize(1 / Lf), tol=1e-9, maxit=12)
    assert it2 <= 3 and np.all(np.isfinite(x2)) and np.all((x2 == -0.2) | (x2 == 0.5) | ((x2 > -0.2) & (x2 < 0.5)))
