"""The numpy oracle (oracle/adaprox_oracle.py) -- the checker of every GPU parity test -- against a second restatement of
src/AdaProx.jl:312-364 written independently in plain C (oracle/adaprox_ref.c): no shared code, no BLAS, sequential
summation.  The two must agree on the stepsize / residual / objective sequences to rounding-level tolerances over a prefix
(the trajectories are chaotic w.r.t. rounding afterwards), and on the final result."""
import numpy as np
import pytest

import adaprox_b200
from oracle import adaprox_oracle as O
from oracle import c_ref as R


def _prefix(log, hist, K, rtol_gamma=1e-10, rtol_res=1e-8, rtol_obj=1e-10):
    K = min(K, len(log), len(hist["gamma"]))
    assert K >= 5
    assert np.allclose([r["gamma"] for r in log[:K]], hist["gamma"][:K], rtol=rtol_gamma, atol=0)
    assert np.allclose([r["sigma"] for r in log[:K]], hist["sigma"][:K], rtol=rtol_gamma, atol=0)
    assert np.allclose([r["norm_res"] for r in log[:K]], hist["norm_res"][:K], rtol=rtol_res, atol=1e-14)
    obj = np.array([r["objective"] for r in log[:K]])
    fin = np.isfinite(obj)
    assert np.array_equal(fin, np.isfinite(hist["objective"][:K]))
    assert np.allclose(obj[fin], hist["objective"][:K][fin], rtol=rtol_obj, atol=1e-13)


@pytest.mark.parametrize("rule", ["our", "mm", "fixed", "plus"])
def test_adapgm_lasso(rule):
    P = adaprox_b200.synth.planted_lasso(100, 300, 10, 0)
    Lf = adaprox_b200.synth.spectral_norm_sq(P["A"], iters=1000, tol=1e-15)
    ro = {"our": O.OurRule(gamma=1 / Lf), "mm": O.MalitskyMishchenkoRule(gamma=1 / Lf), "fixed": O.FixedStepsize(1 / Lf),
          "plus": O.OurRulePlus(gamma=1 / Lf)}[rule]                # defined in src/AdaProx.jl:277-308, used by no experiment
    rc = {"our": R.RULE_OUR, "mm": R.RULE_MM, "fixed": R.RULE_FIXED, "plus": R.RULE_OUR_PLUS}[rule]
    log = []
    xo, ito = O.adaptive_proxgrad(np.zeros(300), f=O.LinearLeastSquares(P["A"], P["b"]), g=O.NormL1(1.0), rule=ro, tol=1e-7, maxit=3000, log=log)
    xc, _, itc, hist = R.adaptive_primal_dual(np.zeros(300), None, f_kind=R.F_LEAST_SQUARES, F=P["A"], fvec=P["b"], g=R.prox_desc(R.P_NORM_L1, 1.0),
                                              rule=rc, gamma=1 / Lf, tol=1e-7, maxit=3000, nhist=3000)
    _prefix(log, hist, 40)
    assert abs(itc - ito) <= max(3, 0.05 * ito)
    fo = O.LinearLeastSquares(P["A"], P["b"])
    oo, oc = fo(xo) + np.abs(xo).sum(), fo(xc) + np.abs(xc).sum()
    assert abs(oo - oc) <= 1e-10 * abs(oo)
    if rule != "fixed":                       # fixed-step PGM is still 1e-6 away after 3000 iterations
        assert abs(oc - P["optimum"]) <= 1e-8 * P["optimum"]


def test_adapgm_dense_logistic():
    rng = np.random.default_rng(1)
    X = rng.standard_normal((150, 20))
    y = (rng.random(150) < 0.5).astype(float)
    gam = 4 * 150 / (np.sum(X * X) + 150)
    log = []
    xo, ito = O.adaptive_proxgrad(np.zeros(21), f=O.LogisticLoss(X, y), g=O.NormL1(0.01), rule=O.OurRule(gamma=gam), tol=1e-8, maxit=3000, log=log)
    xc, _, itc, hist = R.adaptive_primal_dual(np.zeros(21), None, f_kind=R.F_LOGISTIC, F=X, fvec=y, g=R.prox_desc(R.P_NORM_L1, 0.01),
                                              rule=R.RULE_OUR, gamma=gam, tol=1e-8, maxit=3000, nhist=3000)
    _prefix(log, hist, 12)                                         # identical to 1e-14 here; near convergence (24 iterations) the
    _prefix(log, hist, 40, rtol_gamma=1e-8, rtol_res=1e-6)         # differences dx, dgrad cancel and rounding is amplified to ~1e-10
    assert abs(itc - ito) <= max(3, 0.05 * ito) and np.linalg.norm(xo - xc) <= 1e-6 * max(1.0, np.linalg.norm(xo))


@pytest.mark.parametrize("t", [0.1, 1.0])
def test_adapdm_dual_svm(t):
    """dual_svm/runme.jl:47-59: f = Quadratic(Q, q), g = IndBox(0, C), h = IndZero(), A = y'."""
    X, y = adaprox_b200.synth.dense_classification(120, 8, 0)
    Z = y[:, None] * X
    Q, q, N = Z @ Z.T, -np.ones(120), 120
    A = y[None, :].copy()
    nA = float(np.linalg.norm(A))
    log = []
    xo, yo, ito = O.adaptive_primal_dual(np.zeros(N), np.zeros(1), f=O.Quadratic(Q, q), g=O.IndBox(0.0, 0.1), h=O.IndZero(), A=A,
                                         rule=O.OurRule(t=t, norm_A=nA), tol=1e-6, maxit=5000, log=log)
    xc, yc, itc, hist = R.adaptive_primal_dual(np.zeros(N), np.zeros(1), f_kind=R.F_QUADRATIC, F=Q, fvec=q, g=R.prox_desc(R.P_IND_BOX, lo=0.0, hi=0.1),
                                               h=R.prox_desc(R.P_IND_ZERO), A=A, rule=R.RULE_OUR, gamma=1 / (2 * 1.2 * t * nA), t=t, norm_A=nA,
                                               tol=1e-6, maxit=5000, nhist=5000)
    _prefix(log, hist, 40)
    assert abs(itc - ito) <= max(3, 0.05 * ito)
    fq = O.Quadratic(Q, q)
    assert abs(fq(xo) - fq(xc)) <= 1e-8 * abs(fq(xo)) and abs(y @ xc) < 1e-4


@pytest.mark.parametrize("hname", ["l1", "l2"])
def test_adapdm_lad_and_sqrt_lasso(hname):
    """least_absolute_deviation/runme.jl:39-48 and square_root_lasso/runme.jl:41: f = Zero, h = Translate(NormL1 | NormL2, -b)."""
    rng = np.random.default_rng(2)
    A = np.hstack([rng.standard_normal((80, 6)), np.ones((80, 1))])
    b = A @ rng.standard_normal(7) + rng.laplace(size=80)
    nA = float(np.linalg.norm(A))
    ho = O.Translate(O.NormL1() if hname == "l1" else O.NormL2(), -b)
    hc = R.prox_desc(R.P_NORM_L1 if hname == "l1" else R.P_NORM_L2, 1.0, shift=-b)
    log = []
    xo, yo, ito = O.adaptive_primal_dual(np.zeros(7), np.zeros(80), f=O.Zero(), g=O.NormL1(0.5), h=ho, A=A, rule=O.OurRule(t=1.0, norm_A=nA),
                                         tol=1e-6, maxit=4000, log=log)
    xc, yc, itc, hist = R.adaptive_primal_dual(np.zeros(7), np.zeros(80), f_kind=R.F_ZERO, g=R.prox_desc(R.P_NORM_L1, 0.5), h=hc, A=A,
                                               rule=R.RULE_OUR, gamma=1 / (2 * 1.2 * nA), t=1.0, norm_A=nA, tol=1e-6, maxit=4000, nhist=4000)
    _prefix(log, hist, 30)
    assert abs(itc - ito) <= max(3, 0.05 * ito)
    oo = log[min(len(log), ito) - 1]["objective"]
    oc = hist["objective"][min(len(hist["objective"]), itc) - 1]
    assert abs(oo - oc) <= 1e-7 * abs(oo)


def test_condat_vu_fixed_rule():
    """condat_vu (:367-416) = the generic loop with FixedStepsize(gamma, sqrt(sigma / gamma))."""
    X, y = adaprox_b200.synth.dense_classification(60, 5, 1)
    Z = y[:, None] * X
    Q, q, N = Z @ Z.T, -np.ones(60), 60
    A = y[None, :].copy()
    Lf, nA = float(np.linalg.norm(Q)), float(np.linalg.norm(A))
    log = []
    xo, yo, ito = O.condat_vu(np.zeros(N), np.zeros(1), f=O.Quadratic(Q, q), g=O.IndBox(0.0, 1.0), h=O.IndZero(), A=A, Lf=Lf, norm_A=nA,
                              tol=1e-6, maxit=400, log=log)
    alpha = 1.0 if nA > 5 * Lf else 100 * nA / Lf                                # :398-412
    gamma, sigma = 1 / (Lf / 2 + nA / alpha), 0.99 / (nA * alpha)
    xc, yc, itc, hist = R.adaptive_primal_dual(np.zeros(N), np.zeros(1), f_kind=R.F_QUADRATIC, F=Q, fvec=q, g=R.prox_desc(R.P_IND_BOX, lo=0.0, hi=1.0),
                                               h=R.prox_desc(R.P_IND_ZERO), A=A, rule=R.RULE_FIXED, gamma=gamma, t=np.sqrt(sigma / gamma),
                                               tol=1e-6, maxit=400, nhist=400)
    assert itc == ito
    _prefix(log, hist, 400, rtol_res=1e-7)
    assert np.allclose(xo, xc, rtol=0, atol=1e-9) and np.allclose(yo, yc, rtol=0, atol=1e-9)


# ---------------------------------------------------------------- AdaPDM+ (:463-550)
@pytest.mark.parametrize("hname", ["l1", "l2"])
def test_adapdm_plus(hname):
    rng = np.random.default_rng(3)
    A = np.hstack([rng.standard_normal((90, 7)), np.ones((90, 1))])
    b = A @ rng.standard_normal(8) + rng.laplace(size=90)
    nA = float(np.linalg.norm(A))
    ho = O.Translate(O.NormL1() if hname == "l1" else O.NormL2(), -b)
    hc = R.prox_desc(R.P_NORM_L1 if hname == "l1" else R.P_NORM_L2, 1.0, shift=-b)
    for eta in (nA, 0.05 * nA):                                     # eta too small on purpose: the linesearch must grow it
        Ao = O.Counting(A)
        log = []
        xo, yo, ito = O.adaptive_linesearch_primal_dual(np.zeros(8), np.zeros(90), f=O.Zero(), g=O.NormL1(0.5), h=ho, A=Ao, eta=eta, t=1.0,
                                                        tol=1e-6, maxit=3000, log=log)
        xc, yc, itc, hist, trials = R.adaptive_linesearch_primal_dual(np.zeros(8), np.zeros(90), f_kind=R.F_ZERO, g=R.prox_desc(R.P_NORM_L1, 0.5), h=hc,
                                                                      A=A, eta=eta, t=1.0, tol=1e-6, maxit=3000, nhist=3000)
        _prefix(log, hist, 30)
        assert abs(itc - ito) <= max(3, 0.05 * ito)
        if itc == ito:                                              # A' is applied once in the prologue and once per linesearch trial
            assert Ao.amul_count == 1 + trials


# ---------------------------------------------------------------- baselines (:34-192)
def _lasso():
    P = adaprox_b200.synth.planted_lasso(80, 200, 10, 1)
    Lf = adaprox_b200.synth.spectral_norm_sq(P["A"], iters=1000, tol=1e-15)
    return P, Lf


def _prefix3(log, hist, K):
    K = min(K, len(log), len(hist["gamma"]))
    assert K >= 5
    assert np.allclose([r["gamma"] for r in log[:K]], hist["gamma"][:K], rtol=1e-12, atol=0)
    assert np.allclose([r["norm_res"] for r in log[:K]], hist["norm_res"][:K], rtol=1e-8, atol=1e-14)
    assert np.allclose([r["objective"] for r in log[:K]], hist["objective"][:K], rtol=1e-10, atol=0)


@pytest.mark.parametrize("xi", [1.0, 1.5])
def test_backtracking_proxgrad(xi):
    P, Lf = _lasso()
    fo = O.Counting(O.LinearLeastSquares(P["A"], P["b"]))
    log = []
    xo, ito = O.backtracking_proxgrad(np.zeros(200), f=fo, g=O.NormL1(1.0), gamma0=5.0 / Lf, xi=xi, tol=1e-7, maxit=1500, log=log)
    xc, itc, hist, ev = R.proxgrad_family(R.BACKTRACKING_PROXGRAD, np.zeros(200), f_kind=R.F_LEAST_SQUARES, F=P["A"], fvec=P["b"],
                                          g=R.prox_desc(R.P_NORM_L1, 1.0), gamma=5.0 / Lf, xi=xi, tol=1e-7, maxit=1500, nhist=1500)
    _prefix3(log, hist, 40)
    # with xi > 1 every iteration ends on an accept/reject decision `f_z > ub_z` taken near equality: one flipped decision
    # (rounding) shifts the rest of the run, so only the prefix, the optimum and a loose iteration count are comparable
    assert abs(itc - ito) <= max(3, (0.05 if xi == 1.0 else 0.15) * ito)
    fl = O.LinearLeastSquares(P["A"], P["b"])
    assert abs((fl(xo) + np.abs(xo).sum()) - (fl(xc) + np.abs(xc).sum())) <= 1e-10 * P["optimum"]
    if itc == ito:
        assert ev == (fo.eval_count, fo.grad_count)


def test_backtracking_nesterov_and_fixed_nesterov():
    P, Lf = _lasso()
    fo = O.Counting(O.LinearLeastSquares(P["A"], P["b"]))
    log = []
    xo, ito = O.backtracking_nesterov(np.zeros(200), f=fo, g=O.NormL1(1.0), gamma0=5.0 / Lf, tol=1e-7, maxit=1500, log=log)
    xc, itc, hist, ev = R.proxgrad_family(R.BACKTRACKING_NESTEROV, np.zeros(200), f_kind=R.F_LEAST_SQUARES, F=P["A"], fvec=P["b"],
                                          g=R.prox_desc(R.P_NORM_L1, 1.0), gamma=5.0 / Lf, tol=1e-7, maxit=1500, nhist=1500)
    _prefix3(log, hist, 40)
    assert abs(itc - ito) <= max(3, 0.05 * ito)
    if itc == ito:
        assert ev == (fo.eval_count, fo.grad_count)
    for muf in (0.0, 1e-3):                                         # both branches of the (theta, beta) recursion :122-128
        fo = O.Counting(O.LinearLeastSquares(P["A"], P["b"]))
        log = []
        xo, ito = O.fixed_nesterov(np.zeros(200), f=fo, g=O.NormL1(1.0), gamma=1.0 / Lf, muf=muf, tol=1e-7, maxit=1500, log=log)
        xc, itc, hist, ev = R.proxgrad_family(R.FIXED_NESTEROV, np.zeros(200), f_kind=R.F_LEAST_SQUARES, F=P["A"], fvec=P["b"],
                                              g=R.prox_desc(R.P_NORM_L1, 1.0), gamma=1.0 / Lf, muf=muf, tol=1e-7, maxit=1500, nhist=1500)
        _prefix3(log, hist, 40)
        assert abs(itc - ito) <= max(3, 0.05 * ito)
        if itc == ito:
            assert ev == (fo.eval_count, fo.grad_count)


@pytest.mark.parametrize("gamma0", [None, 0.7])
def test_agraal(gamma0):
    P, Lf = _lasso()
    x_second = np.random.default_rng(4).standard_normal(200)
    g0 = None if gamma0 is None else gamma0 / Lf
    fo = O.Counting(O.LinearLeastSquares(P["A"], P["b"]))
    log = []
    xo, ito = O.agraal(np.zeros(200), f=fo, g=O.NormL1(1.0), x0=x_second, gamma0=g0, tol=1e-7, maxit=1500, log=log)
    xc, itc, hist, ev = R.proxgrad_family(R.AGRAAL, np.zeros(200), f_kind=R.F_LEAST_SQUARES, F=P["A"], fvec=P["b"], g=R.prox_desc(R.P_NORM_L1, 1.0),
                                          gamma=0.0 if g0 is None else g0, x_second=x_second, tol=1e-7, maxit=1500, nhist=1500)
    K = min(30, len(log), len(hist["gamma"]))
    assert np.allclose([r["gamma"] for r in log[:K]], hist["gamma"][:K], rtol=1e-10, atol=0)
    assert np.allclose([r["objective"] for r in log[:K]], hist["objective"][:K], rtol=1e-10, atol=0)
    assert abs(itc - ito) <= max(3, 0.05 * ito)
    if itc == ito:
        assert ev == (fo.eval_count, fo.grad_count)


# ---------------------------------------------------------------- Malitsky-Pock (:555-629)
@pytest.mark.parametrize("t", [0.5, 2.0])
def test_malitsky_pock(t):
    X, y = adaprox_b200.synth.dense_classification(100, 7, 2)
    Z = y[:, None] * X
    Q, q, N = Z @ Z.T, -np.ones(100), 100
    A = y[None, :].copy()
    nA = float(np.linalg.norm(A))
    log = []
    xo, yo, ito = O.malitsky_pock(np.zeros(N), np.zeros(1), f=O.Quadratic(Q, q), g=O.IndBox(0.0, 0.1), h=O.IndZero(), A=A, sigma=1 / nA, t=t,
                                  tol=1e-6, maxit=600, log=log)
    xc, yc, itc, hist = R.malitsky_pock(np.zeros(N), np.zeros(1), f_kind=R.F_QUADRATIC, F=Q, fvec=q, g=R.prox_desc(R.P_IND_BOX, lo=0.0, hi=0.1),
                                        h=R.prox_desc(R.P_IND_ZERO), A=A, sigma=1 / nA, t=t, tol=1e-6, maxit=600, nhist=600)
    _prefix(log, hist, 40)
    assert abs(itc - ito) <= max(3, 0.05 * ito)


# ---------------------------------------------------------------- prox objects, both restatements
def test_prox_objects_agree():
    """Every prox object x {plain, Translate} x {itself, convex_conjugate}: the numpy oracle and the C restatement apply the same
    operations in the same order (Moreau in ProximalCore's order), so they agree to the last bits on random inputs."""
    rng = np.random.default_rng(5)
    n = 257
    shift = rng.standard_normal(n)
    kinds = [("zero", lambda: O.Zero(), lambda s: R.prox_desc(R.P_ZERO, shift=s)),
             ("indzero", lambda: O.IndZero(), lambda s: R.prox_desc(R.P_IND_ZERO, shift=s)),
             ("l1", lambda: O.NormL1(0.7), lambda s: R.prox_desc(R.P_NORM_L1, 0.7, shift=s)),
             ("l2", lambda: O.NormL2(1.3), lambda s: R.prox_desc(R.P_NORM_L2, 1.3, shift=s)),
             ("box", lambda: O.IndBox(-0.4, 0.9), lambda s: R.prox_desc(R.P_IND_BOX, lo=-0.4, hi=0.9, shift=s))]
    for name, mk_o, mk_c in kinds:
        for translated in (False, True):
            fo = O.Translate(mk_o(), shift) if translated else mk_o()
            fc = mk_c(shift if translated else None)
            for gamma in (0.05, 1.0, 7.5):
                for scale in (0.1, 3.0):
                    x = scale * rng.standard_normal(n)
                    yo, vo = O.prox(fo, x, gamma)
                    yc, vc = R.prox_eval(fc, x, gamma)
                    assert np.allclose(yo, yc, rtol=0, atol=1e-15 * max(1.0, np.max(np.abs(x)))), (name, translated, gamma)
                    # the C entry point reports f evaluated AT y (what the loops log as g(x) / h(A x)); for a translated indicator
                    # that may differ from the value prox returns (computed before `y .-= b`) by one rounding of (v - b) + b
                    vo = fo(yc)
                    assert (np.isinf(vo) and np.isinf(vc)) or abs(vo - vc) <= 1e-13 * max(1.0, abs(vo)), (name, translated, gamma)
                    yo2, _ = O.prox(O.convex_conjugate(fo), x, gamma)
                    yc2, _ = R.prox_eval(fc, x, gamma, conjugate=True)
                    assert np.allclose(yo2, yc2, rtol=0, atol=4e-15 * max(1.0, np.max(np.abs(x)))), (name, translated, gamma, "conj")


# ---------------------------------------------------------------- seeded random combinations of f, g, h, A and rule
def test_random_combinations_of_the_generic_loop():
    """40 seeded instances of adaptive_primal_dual with random sizes and a random choice of smooth term, prox objects, linear map
    and stepsize rule: the first 15 stepsizes / residuals of the two restatements agree (rounding only)."""
    rng = np.random.default_rng(2024)
    for case in range(40):
        n, md = int(rng.integers(3, 40)), int(rng.integers(2, 30))
        fk = rng.choice(["ls", "quad", "zero", "logistic"])
        if fk == "ls":
            m = int(rng.integers(2, 50)); F = rng.standard_normal((m, n)); fv = rng.standard_normal(m)
            fo, fc, Lf = O.LinearLeastSquares(F, fv), dict(f_kind=R.F_LEAST_SQUARES, F=F, fvec=fv), np.linalg.norm(F, 2) ** 2
        elif fk == "quad":
            B = rng.standard_normal((n, n)); F = B @ B.T / n; fv = rng.standard_normal(n)
            fo, fc, Lf = O.Quadratic(F, fv), dict(f_kind=R.F_QUADRATIC, F=F, fvec=fv), np.linalg.norm(F, 2)
        elif fk == "logistic":
            m = int(rng.integers(5, 60)); F = rng.standard_normal((m, n - 1)); fv = (rng.random(m) < 0.5).astype(float)
            fo, fc, Lf = O.LogisticLoss(F, fv), dict(f_kind=R.F_LOGISTIC, F=F, fvec=fv), (np.sum(F * F) + m) / (4 * m)
        else:
            fo, fc, Lf = O.Zero(), dict(f_kind=R.F_ZERO), 0.0
        gk = rng.choice(["l1", "box", "zero"])
        go, gc = {"l1": (O.NormL1(0.3), R.prox_desc(R.P_NORM_L1, 0.3)), "box": (O.IndBox(-0.5, 0.8), R.prox_desc(R.P_IND_BOX, lo=-0.5, hi=0.8)),
                  "zero": (O.Zero(), R.prox_desc(R.P_ZERO))}[gk]
        A = rng.standard_normal((md, n))
        b = rng.standard_normal(md)
        hk = rng.choice(["l1t", "l2t", "indzero", "l1"])
        ho, hc = {"l1t": (O.Translate(O.NormL1(), -b), R.prox_desc(R.P_NORM_L1, 1.0, shift=-b)),
                  "l2t": (O.Translate(O.NormL2(), -b), R.prox_desc(R.P_NORM_L2, 1.0, shift=-b)),
                  "indzero": (O.IndZero(), R.prox_desc(R.P_IND_ZERO)), "l1": (O.NormL1(0.5), R.prox_desc(R.P_NORM_L1, 0.5))}[hk]
        nA = float(np.linalg.norm(A))
        t = float(rng.choice([0.3, 1.0, 2.5]))
        rk = rng.choice(["our", "mm", "fixed"])
        if rk == "our":
            ro, rc = O.OurRule(t=t, norm_A=nA), dict(rule=R.RULE_OUR, gamma=1 / (2 * 1.2 * t * nA), t=t, norm_A=nA)
        else:
            gam = 0.5 / (Lf + nA * max(t, 1.0) + 1e-3)
            ro = O.MalitskyMishchenkoRule(gamma=gam, t=t) if rk == "mm" else O.FixedStepsize(gam, t)
            rc = dict(rule=R.RULE_MM if rk == "mm" else R.RULE_FIXED, gamma=gam, t=t)
        x0, y0 = rng.standard_normal(n), rng.standard_normal(md)
        log = []
        O.adaptive_primal_dual(x0, y0, f=fo, g=go, h=ho, A=A, rule=ro, tol=1e-9, maxit=60, log=log)
        xc, yc, itc, hist = R.adaptive_primal_dual(x0, y0, g=gc, h=hc, A=A, tol=1e-9, maxit=60, nhist=60, **fc, **rc)
        K = min(15, len(log), len(hist["gamma"]))
        res0 = np.array([r["norm_res"] for r in log[:K]])
        grow = np.nonzero(res0 > 10 * res0[0])[0]                  # an unstable (gamma, t) choice: the run diverges and amplifies
        if len(grow):                                              # rounding exponentially -- compare up to that point only
            K = max(4, int(grow[0]))
        tag = (case, fk, gk, hk, rk, t)
        assert K >= 1, tag
        go_, gc_ = np.array([r["gamma"] for r in log[:K]]), hist["gamma"][:K]
        both_nan = np.isnan(go_) & np.isnan(gc_)
        assert np.allclose(go_[~both_nan], gc_[~both_nan], rtol=1e-9, atol=0), tag
        ro_, rc_ = np.array([r["norm_res"] for r in log[:K]]), hist["norm_res"][:K]
        fin = np.isfinite(ro_) & np.isfinite(rc_)
        assert np.array_equal(np.isfinite(ro_), np.isfinite(rc_)), tag
        assert np.allclose(ro_[fin], rc_[fin], rtol=1e-7, atol=1e-12), tag


def test_nan_stepsize_takes_the_same_course_in_both_restatements():
    """The degenerate AdaPGM instance of test_oracle_known_answers (Malitsky-Mishchenko, box keeps the iterate in place: gamma = NaN from
    the second iteration on): both restatements carry the NaN through w = y + sigma * (...) * 0 into norm_res and x, neither stops."""
    m, n = 149, 3
    rng = np.random.default_rng(m * 31 + n)
    A = np.asfortranarray(rng.standard_normal((m, n)) / np.sqrt(m))
    b = rng.standard_normal(m)
    Lf = float(np.linalg.norm(A, 2) ** 2)
    rng.standard_normal(n)
    x0 = 0.05 * rng.standard_normal(n)
    log = []
    xo, ito = O.adaptive_proxgrad(x0, f=O.LinearLeastSquares(A, b), g=O.IndBox(-0.2, 0.5), rule=O.MalitskyMishchenkoRule(gamma=1 / Lf),
                                  tol=1e-9, maxit=12, log=log)
    xc, _, itc, hist = R.adaptive_primal_dual(x0, None, f_kind=R.F_LEAST_SQUARES, F=A, fvec=b, g=R.prox_desc(R.P_IND_BOX, lo=-0.2, hi=0.5),
                                              rule=R.RULE_MM, gamma=1 / Lf, tol=1e-9, maxit=12, nhist=12)
    assert ito == itc == 12 and np.all(np.isnan(xo)) and np.all(np.isnan(xc))
    for key in ("gamma", "norm_res"):
        a = np.array([r[key] for r in log])
        assert np.array_equal(np.isnan(a), np.isnan(hist[key][:12])) and np.isnan(a[1:]).all() and np.isfinite(a[0])
        assert abs(a[0] - hist[key][0]) <= 1e-12 * abs(a[0])
