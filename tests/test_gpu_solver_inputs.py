"""Solver inputs added in round 2, through the C ABI against the CPU oracle: g = NormL2 and a conjugate g in the adaptive
loops (src/AdaProx.jl:332,361 accept any prox-able g), h passed as a conjugate, auto_adaptive_proxgrad (:423-455), the Cubic
oracle as a SOLVER input with its logistic_loss_grad_Hessian setup (cubic_sparse_logreg/runme.jl:20-45,66-120)."""
import numpy as np
import pytest

from oracle import adaprox_oracle as O
from oracle import drift

pytestmark = pytest.mark.gpu


def _gam(log, k=None):
    return np.array([r["gamma"] for r in (log if k is None else log[:k])])


def _ls(AdaProx, m=80, n=120, seed=3):
    rng = np.random.default_rng(seed)
    A = np.asfortranarray(rng.standard_normal((m, n)) / np.sqrt(m))
    b = rng.standard_normal(m)
    Lf = float(np.linalg.norm(A, 2) ** 2)
    return A, b, Lf


@pytest.mark.parametrize("gname", ["l2", "l2_translated", "conj_l1", "conj_l2", "conj_box"])
def test_adapgm_with_norml2_and_conjugate_g(AdaProx, gname):
    A, b, Lf = _ls(AdaProx)
    n = A.shape[1]
    rng = np.random.default_rng(5)
    c = 0.3 * rng.standard_normal(n)
    gd, go = {
        "l2": (AdaProx.NormL2(0.8), O.NormL2(0.8)),
        "l2_translated": (AdaProx.Translate(AdaProx.NormL2(0.5), -c), O.Translate(O.NormL2(0.5), -c)),
        "conj_l1": (AdaProx.convex_conjugate(AdaProx.NormL1(0.7)), O.convex_conjugate(O.NormL1(0.7))),        # = indicator of the 0.7-box
        "conj_l2": (AdaProx.convex_conjugate(AdaProx.NormL2(1.5)), O.convex_conjugate(O.NormL2(1.5))),        # = indicator of the 1.5-ball
        "conj_box": (AdaProx.convex_conjugate(AdaProx.IndBox(-0.2, 0.4)), O.convex_conjugate(O.IndBox(-0.2, 0.4))),   # support function of the box
    }[gname]
    x0 = 0.1 * rng.standard_normal(n)
    fd, fo = AdaProx.Counting(AdaProx.LinearLeastSquares(A, b)), O.Counting(O.LinearLeastSquares(A, b))
    # a logger would call g(x), which a ConvexConjugate object does not support in the reference: records without objective
    xd, itd = AdaProx.adaptive_proxgrad(x0, f=fd, g=gd, rule=AdaProx.OurRule(gamma=1 / Lf), tol=1e-8, maxit=4000)
    xo, ito = O.adaptive_proxgrad(x0, f=fo, g=go, rule=O.OurRule(gamma=1 / Lf), tol=1e-8, maxit=4000)
    assert abs(itd - ito) <= max(3, 0.05 * ito), (itd, ito)
    assert np.linalg.norm(xd - xo) <= 1e-6 * max(np.linalg.norm(xo), 1.0)
    assert fd.eval_count == itd + 1 and fd.grad_count == itd + 1
    # teacher-free prefix with records where g is callable
    if not gname.startswith("conj"):
        logd, logo = [], []
        AdaProx.adaptive_proxgrad(x0, f=fd, g=gd, rule=AdaProx.OurRule(gamma=1 / Lf), tol=0.0, maxit=25, log=logd)
        O.adaptive_proxgrad(x0, f=fo, g=go, rule=O.OurRule(gamma=1 / Lf), tol=0.0, maxit=25, log=logo)
        assert np.max(np.abs(_gam(logd) / _gam(logo) - 1)) < 1e-11
        assert np.allclose([r["objective"] for r in logd], [r["objective"] for r in logo], rtol=1e-11)
        assert np.allclose([r["norm_res"] for r in logd], [r["norm_res"] for r in logo], rtol=1e-9)
    # the comparison baselines have no reduction before their prox: refused, never a silent wrong answer
    with pytest.raises(AdaProx.AdaproxError):
        AdaProx.backtracking_proxgrad(x0, f=fd, g=gd, gamma0=1 / Lf, maxit=5)


def test_adapdm_with_norml2_g_and_conjugate_h(AdaProx):
    """AdaPDM with g = NormL2 (one more reduction in the primal step) and h handed over as a conjugate: h = (lam |.|_1)* is the
    indicator of the lam-box, so convex_conjugate(h) is NormL1 again and the dual step is its plain prox."""
    rng = np.random.default_rng(8)
    m, n = 50, 70
    A = rng.standard_normal((m, n)) / np.sqrt(n)
    F = np.asfortranarray(rng.standard_normal((40, n)) / np.sqrt(40)); bf = rng.standard_normal(40)
    nA = float(np.linalg.norm(A, 2))
    hd = AdaProx.convex_conjugate(AdaProx.NormL1(0.6)); ho = O.convex_conjugate(O.NormL1(0.6))
    kw = dict(tol=0.0, maxit=40)
    logd = []
    xd, yd, itd = AdaProx.adaptive_primal_dual(np.zeros(n), np.zeros(m), f=AdaProx.LinearLeastSquares(F, bf), g=AdaProx.NormL2(0.3), h=hd,
                                              A=AdaProx.DeviceMatrix(A), rule=AdaProx.OurRule(t=1.0, norm_A=nA), **kw)
    xo, yo, ito = O.adaptive_primal_dual(np.zeros(n), np.zeros(m), f=O.LinearLeastSquares(F, bf), g=O.NormL2(0.3), h=ho, A=A,
                                         rule=O.OurRule(t=1.0, norm_A=nA), **kw)
    assert itd == ito == 40
    assert np.linalg.norm(xd - xo) <= 1e-10 * np.linalg.norm(xo) and np.linalg.norm(yd - yo) <= 1e-10 * max(np.linalg.norm(yo), 1e-300)
    # the dual iterate of h* = lam |.|_1 ... its prox is the soft threshold: y is sparse
    assert np.count_nonzero(yd) == np.count_nonzero(yo)


def test_auto_adaptive_proxgrad(AdaProx, lasso_small):
    P = lasso_small
    n = 1000
    for gamma in (1.0 / P["Lf"], 50.0 / P["Lf"], 1e7 / P["Lf"]):       # the last one triggers the "initial guess too large" branch (:445-450)
        logd, logo = [], []
        xd, itd = AdaProx.auto_adaptive_proxgrad(np.zeros(n), f=AdaProx.LinearLeastSquares(P["A"], P["b"]), g=AdaProx.NormL1(1.0), gamma=gamma,
                                                 tol=1e-6, maxit=10_000, log=logd)
        xo, ito = O.auto_adaptive_proxgrad(np.zeros(n), f=O.LinearLeastSquares(P["A"], P["b"]), g=O.NormL1(1.0), gamma=gamma, tol=1e-6,
                                           maxit=10_000, log=logo)
        assert abs(logd[0]["gamma"] / logo[0]["gamma"] - 1) < 1e-12       # the estimated initial stepsize
        assert np.max(np.abs(_gam(logd, 15) / _gam(logo, 15) - 1)) < 1e-11
        assert abs(itd - ito) <= max(3, 0.05 * ito)
        assert abs(logd[-1]["objective"] - logo[-1]["objective"]) <= 1e-10 * abs(logo[-1]["objective"])
    x, it = AdaProx.auto_adaptive_proxgrad(P["x_star"] * 0, f=AdaProx.LinearLeastSquares(P["A"], 0 * P["b"]), g=AdaProx.NormL1(1.0), gamma=1.0)
    assert it == 0                                                      # zero gradient at the start: returns at once (:426-428)
    with pytest.raises(TypeError):
        AdaProx.auto_adaptive_proxgrad(np.zeros(n), f=AdaProx.LinearLeastSquares(P["A"], P["b"]), g=AdaProx.NormL1(1.0))


@pytest.mark.parametrize("sparse", [False, True])
def test_logistic_loss_grad_hessian(AdaProx, sparse):
    import scipy.sparse as sp
    rng = np.random.default_rng(2)
    m, n = 300, 40
    X = sp.random(m, n, density=0.2, random_state=4, format="csr") if sparse else rng.standard_normal((m, n))
    y = (rng.random(m) < 0.4).astype(float)
    for w in (np.zeros(n + 1), 0.3 * rng.standard_normal(n + 1)):
        Hd, gd = AdaProx.logistic_loss_grad_Hessian(X, y, w)
        Ho, go = O.logistic_loss_grad_Hessian(X, y, w)
        assert Hd.shape == (n + 1, n + 1) and np.max(np.abs(Hd - Ho)) <= 1e-13 * np.max(np.abs(Ho))
        assert np.max(np.abs(Hd - Hd.T)) <= 1e-15 * np.max(np.abs(Hd))
        assert np.max(np.abs(gd - go)) <= 1e-13 * np.max(np.abs(go))


def test_cubic_subproblem_solvers(AdaProx):
    """cubic_sparse_logreg/runme.jl:47-160 on a synthetic data set of the mushrooms shape class: Cubic(Q, q, lam) built from the
    logistic Hessian at 0, g = Zero, gam_init from a random perturbation, then every solver the script runs."""
    import scipy.sparse as sp
    rng = np.random.default_rng(0)
    m, n = 600, 60
    X = sp.random(m, n, density=0.25, random_state=1, format="csr", data_rvs=lambda k: rng.random(k))
    y = (rng.random(m) < 0.5).astype(float)
    n1 = n + 1
    Qd, qd = AdaProx.logistic_loss_grad_Hessian(X, y, np.zeros(n1))             # :64
    Qo, qo = O.logistic_loss_grad_Hessian(X, y, np.zeros(n1))
    assert np.max(np.abs(Qd - Qo)) <= 1e-13 * np.max(np.abs(Qo))
    lam = 1.0
    fd_raw, fo_raw = AdaProx.Cubic(Qd, qd, lam), O.Cubic(Qo, qo, lam)
    x0 = np.zeros(n1)
    x_pert = x0 + rng.standard_normal(n1)                                       # :71
    _, g0 = O.eval_with_gradient(fo_raw, x0); _, gp = O.eval_with_gradient(fo_raw, x_pert)
    gam_init = float(np.linalg.norm(x0 - x_pert) ** 2 / np.dot(g0 - gp, x0 - x_pert))   # :75
    _, g0d = AdaProx.eval_with_gradient(fd_raw, x0); _, gpd = AdaProx.eval_with_gradient(fd_raw, x_pert)
    assert abs(float(np.linalg.norm(x0 - x_pert) ** 2 / np.dot(g0d - gpd, x0 - x_pert)) / gam_init - 1) < 1e-12
    tol, maxit = 1e-5, 1000

    def both(name, **kw):
        fd, fo = AdaProx.Counting(fd_raw), O.Counting(fo_raw)
        logd, logo = [], []
        xd, itd = getattr(AdaProx, name)(x0, f=fd, g=AdaProx.Zero(), tol=tol, maxit=maxit, log=logd, **{k: (v[0] if isinstance(v, tuple) else v) for k, v in kw.items()})
        xo, ito = getattr(O, name)(x0, f=fo, g=O.Zero(), tol=tol, maxit=maxit, log=logo, **{k: (v[1] if isinstance(v, tuple) else v) for k, v in kw.items()})
        K = min(20, len(logd), len(logo))
        assert np.max(np.abs(_gam(logd, K) / _gam(logo, K) - 1)) < 1e-10, name
        assert abs(itd - ito) <= max(2, 0.03 * ito), (name, itd, ito)
        assert abs(logd[-1]["objective"] - logo[-1]["objective"]) <= 1e-10 * max(abs(logo[-1]["objective"]), 1e-3), name
        assert np.linalg.norm(xd - xo) <= 1e-4 * max(np.linalg.norm(xo), 1e-9), name
        if itd == ito:
            assert (fd.eval_count, fd.grad_count) == (fo.eval_count, fo.grad_count), name
        return itd

    both("adaptive_proxgrad", rule=(AdaProx.OurRule(gamma=gam_init), O.OurRule(gamma=gam_init)))                 # :78-86, :122-130
    both("adaptive_proxgrad", rule=(AdaProx.MalitskyMishchenkoRule(gamma=gam_init), O.MalitskyMishchenkoRule(gamma=gam_init)))   # :112-120
    for xi in (1, 1.5, 2):
        both("backtracking_proxgrad", gamma0=gam_init, xi=xi)                                                    # :87-99
    both("backtracking_nesterov", gamma0=gam_init)                                                               # :101-110
    fd, fo = AdaProx.Counting(fd_raw), O.Counting(fo_raw)
    xd, itd = AdaProx.agraal(x0, f=fd, g=AdaProx.Zero(), x0=x_pert, gamma0=gam_init, tol=tol, maxit=maxit)       # :132-142
    xo, ito = O.agraal(x0, f=fo, g=O.Zero(), x0=x_pert, gamma0=gam_init, tol=tol, maxit=maxit)
    assert abs(itd - ito) <= max(2, 0.03 * ito) and np.linalg.norm(xd - xo) <= 1e-4 * max(np.linalg.norm(xo), 1e-9)
    # the minimiser of the cubic model: gradient vanishes
    _, gsol = O.eval_with_gradient(fo_raw, xd)
    assert np.linalg.norm(gsol) <= 10 * tol


# ---------------------------------------------------------------- the cluster-resident small-problem kernel (solver_resident.cuh)
@pytest.mark.parametrize("m,n", [(400, 1000), (100, 300), (7, 1), (5, 1024), (33, 517), (416, 1024)])
def test_resident_kernel_matches_grid_kernel_and_oracle(AdaProx, m, n):
    """Small dense least squares runs with the matrix resident in the shared memory of one 16-CTA cluster.  Shapes: the reference's
    own lasso sizes (lasso/runme.jl:192-207), fewer rows than CTAs, a single column, the widest row (n = 1024), ragged everything,
    and the largest matrix that still fits; every stepsize rule; box and translated-l1 prox; maxit = 0 and 1."""
    _small_ls_kernel_check(AdaProx, m, n, {"ADAPROX_RESIDENT": "1"}, {"ADAPROX_RESIDENT": "0", "ADAPROX_GRIDRES": "0"}, 3)


# ---------------------------------------------------------------- the grid-resident kernel (solver_gridres.cuh)
@pytest.mark.parametrize("m,n", [(500, 1000), (4000, 1000), (120, 40), (1, 1024), (3000, 517), (4100, 1024), (4144, 1000)])
def test_grid_resident_kernel_matches_grid_kernel_and_oracle(AdaProx, m, n):
    """Dense least squares with rows of at most 1024 columns that fits the shared memory of all SMs together: the reference's 500 x 1000 and
    4000 x 1000 lasso runs (lasso/runme.jl:191-195), more CTAs than rows, one row, ragged shapes, the widest row with every CTA full
    (4100 x 1024: 28 rows of 8 KB per CTA), 148 x 28 rows exactly; same checks as the cluster-resident kernel."""
    _small_ls_kernel_check(AdaProx, m, n, {"ADAPROX_RESIDENT": "0", "ADAPROX_GRIDRES": "1"}, {"ADAPROX_RESIDENT": "0", "ADAPROX_GRIDRES": "0"}, 4)


@pytest.mark.parametrize("m,n", [(500, 1000), (120, 40), (4000, 1000)])
def test_grid_resident_nesterov_and_agraal(AdaProx, m, n):
    """fixed_nesterov (src/AdaProx.jl:91-142), agraal (:150-192), backtracking_proxgrad (:50-64) and backtracking_nesterov (:66-84) on a
    small dense least-squares term run in the grid-resident kernel (MODE 1-4 of solver_gridres.cuh).  Against the oracle and against the general grid kernel (ADAPROX_GRIDRES=0): stepsizes,
    residuals, objectives of the records, counters (the logged f(x) is not counted), with and without records, maxit = 0 / 1."""
    import os
    P = AdaProx.synth.planted_lasso(m, n, 10, 0)
    Lf = float(np.linalg.norm(P["A"], 2) ** 2)
    g0 = 1.0 / Lf
    xs0 = np.random.default_rng(3).standard_normal(n)
    calls = {"nesterov": (lambda M, f, g, **kw: M.fixed_nesterov(np.zeros(n), f=f, g=g, gamma=g0, **kw)),
             "nesterov_mu": (lambda M, f, g, **kw: M.fixed_nesterov(np.zeros(n), f=f, g=g, gamma=g0, muf=0.05 * Lf, **kw)),
             "agraal": (lambda M, f, g, **kw: M.agraal(np.zeros(n), f=f, g=g, x0=xs0, gamma0=g0, **kw)),
             "agraal_auto": (lambda M, f, g, **kw: M.agraal(np.zeros(n), f=f, g=g, x0=xs0, **kw)),
             "backtracking_xi1": (lambda M, f, g, **kw: M.backtracking_proxgrad(np.zeros(n), f=f, g=g, gamma0=10 * g0, xi=1.0, **kw)),
             "backtracking_xi2": (lambda M, f, g, **kw: M.backtracking_proxgrad(np.zeros(n), f=f, g=g, gamma0=g0, xi=2.0, **kw)),
             "backtracking_nesterov": (lambda M, f, g, **kw: M.backtracking_nesterov(np.zeros(n), f=f, g=g, gamma0=10 * g0, **kw))}
    fd_raw = AdaProx.LinearLeastSquares(P["A"], P["b"])
    for name, call in calls.items():
        fo = O.Counting(O.LinearLeastSquares(P["A"], P["b"]))
        lo = []
        xo, ito = call(O, fo, O.NormL1(1.0), tol=1e-7, maxit=150, log=lo)
        got = {}
        for mode in ("1", "0"):
            os.environ["ADAPROX_GRIDRES"] = mode
            try:
                fd = AdaProx.Counting(fd_raw)
                ld = []
                xd, itd = call(AdaProx, fd, AdaProx.NormL1(1.0), tol=1e-7, maxit=150, log=ld)
                passes = AdaProx.last_solve_info()["matrix_passes"]
                xq, itq = call(AdaProx, fd_raw, AdaProx.NormL1(1.0), tol=1e-7, maxit=150)          # no records: no value-only evaluations
                got[mode] = (xd, itd, ld, passes, (fd.eval_count, fd.grad_count), xq, itq)
            finally:
                os.environ.pop("ADAPROX_GRIDRES", None)
        assert got["1"][3] == 4 and got["0"][3] == 2, (name, got["1"][3], got["0"][3])
        for mode in ("1", "0"):
            xd, itd, ld, _, counts, xq, itq = got[mode]
            K = min(30, len(ld), len(lo))
            assert abs(itd - ito) <= max(2, 0.03 * ito), (name, mode, itd, ito)
            assert counts == (fo.eval_count, fo.grad_count) or itd != ito, (name, mode, counts)
            assert np.allclose([r["gamma"] for r in ld[:K]], [r["gamma"] for r in lo[:K]], rtol=1e-11), (name, mode)
            assert np.allclose([r["objective"] for r in ld[:K]], [r["objective"] for r in lo[:K]], rtol=1e-10), (name, mode)
            assert np.allclose([r["norm_res"] for r in ld[:K]], [r["norm_res"] for r in lo[:K]], rtol=1e-8), (name, mode)
            assert [r["f_evals"] for r in ld[:K]] == [r["f_evals"] for r in lo[:K]], (name, mode)
            assert [r["grad_f_evals"] for r in ld[:K]] == [r["grad_f_evals"] for r in lo[:K]], (name, mode)
            assert [r["prox_g_evals"] for r in ld[:K]] == [r["prox_g_evals"] for r in lo[:K]], (name, mode)
            assert np.linalg.norm(xd - xo) <= 1e-4 * np.linalg.norm(xo), (name, mode)     # 150 iterations of a free-running trajectory (cf. test_cubic_*: 1e-4)
            assert itq == itd and np.array_equal(xq, xd), (name, mode, "records must not change the iterates")
    for maxit in (0, 1):
        os.environ["ADAPROX_GRIDRES"] = "1"
        try:
            xd, itd = AdaProx.agraal(np.zeros(n), f=fd_raw, g=AdaProx.NormL1(1.0), x0=xs0, gamma0=g0, tol=0.0, maxit=maxit)
            xn, itn = AdaProx.fixed_nesterov(np.zeros(n), f=fd_raw, g=AdaProx.NormL1(1.0), gamma=g0, tol=0.0, maxit=maxit)
        finally:
            os.environ.pop("ADAPROX_GRIDRES", None)
        xo, ito = O.agraal(np.zeros(n), f=O.LinearLeastSquares(P["A"], P["b"]), g=O.NormL1(1.0), x0=xs0, gamma0=g0, tol=0.0, maxit=maxit)
        xno, itno = O.fixed_nesterov(np.zeros(n), f=O.LinearLeastSquares(P["A"], P["b"]), g=O.NormL1(1.0), gamma=g0, tol=0.0, maxit=maxit)
        assert itd == ito == maxit and np.linalg.norm(xd - xo) <= 1e-12 * max(np.linalg.norm(xo), 1e-12)
        assert itn == itno == maxit and np.linalg.norm(xn - xno) <= 1e-12 * max(np.linalg.norm(xno), 1e-12)
    fd_raw.mat.free()


def test_nonfinite_stepsize_reaches_the_residual_like_the_reference(AdaProx):
    """src/AdaProx.jl:342-348 with A = 0, h = Zero: once the rule returns a NaN stepsize (Malitsky-Mishchenko on an iterate the box keeps
    in place: dx = 0, L = 0 / 0), w = y + sigma * (...) * 0 is NaN, so norm_res is NaN, the test :354 never fires and x turns NaN -- although
    the primal residual alone is exactly 0 at that point.  Every AdaPGM kernel has to do the same, not stop with 'converged'."""
    import os
    m, n = 149, 3
    rng = np.random.default_rng(m * 31 + n)
    A = np.asfortranarray(rng.standard_normal((m, n)) / np.sqrt(m))
    b = rng.standard_normal(m)
    Lf = float(np.linalg.norm(A, 2) ** 2)
    rng.standard_normal(n)
    x0 = 0.05 * rng.standard_normal(n)
    logo = []
    xo, ito = O.adaptive_proxgrad(x0, f=O.LinearLeastSquares(A, b), g=O.IndBox(-0.2, 0.5), rule=O.MalitskyMishchenkoRule(gamma=1 / Lf), tol=1e-9, maxit=12, log=logo)
    assert ito == 12 and np.all(np.isnan(xo)) and np.isnan(logo[1]["norm_res"]) and np.isfinite(logo[0]["norm_res"])   # the instance is the degenerate one
    f = AdaProx.LinearLeastSquares(A, b)
    for env, passes in (({"ADAPROX_RESIDENT": "1"}, 3), ({"ADAPROX_RESIDENT": "0", "ADAPROX_GRIDRES": "1"}, 4),
                        ({"ADAPROX_RESIDENT": "0", "ADAPROX_GRIDRES": "0"}, 2), ({"ADAPROX_FUSED": "1"}, 1)):
        os.environ.update(env)
        try:
            log = []
            x, it = AdaProx.adaptive_proxgrad(x0, f=f, g=AdaProx.IndBox(-0.2, 0.5), rule=AdaProx.MalitskyMishchenkoRule(gamma=1 / Lf), tol=1e-9, maxit=12, log=log)
            assert AdaProx.last_solve_info()["matrix_passes"] == passes, env
        finally:
            for k_ in env:
                os.environ.pop(k_, None)
        assert it == 12 and np.all(np.isnan(x)), (env, it, x)
        assert np.array_equal(np.isnan([r["norm_res"] for r in log]), np.isnan([r["norm_res"] for r in logo])), env
        assert np.array_equal(np.isnan([r["gamma"] for r in log]), np.isnan([r["gamma"] for r in logo])), env
        assert abs(log[0]["norm_res"] - logo[0]["norm_res"]) <= 1e-12 * logo[0]["norm_res"]
    f.mat.free()


def _small_ls_kernel_check(AdaProx, m, n, env_on, env_off, passes_on):
    import os
    rng = np.random.default_rng(m * 31 + n)
    A = np.asfortranarray(rng.standard_normal((m, n)) / np.sqrt(max(m, 2)))
    b = rng.standard_normal(m)
    Lf = float(np.linalg.norm(A, 2) ** 2) if min(m, n) > 1 else float(np.sum(A * A))
    c = 0.1 * rng.standard_normal(n)
    x0 = 0.05 * rng.standard_normal(n)
    cases = [("our", AdaProx.NormL1(0.3), lambda pm: O.NormL1(0.3)), ("mm", AdaProx.IndBox(-0.2, 0.5), lambda pm: O.IndBox(-0.2, 0.5)),
             ("fixed", AdaProx.Translate(AdaProx.NormL1(0.2), -c), lambda pm: O.Translate(O.NormL1(0.2), -c[pm])), ("plus", AdaProx.Zero(), lambda pm: O.Zero())]
    f = AdaProx.LinearLeastSquares(A, b)
    gscale = float(np.linalg.norm(A.T @ (A @ x0 - b))) + 1e-300      # residuals are compared down to 1e-12 of the initial gradient
    fscale = 0.5 * float(np.sum((A @ x0 - b) ** 2))                   # ... objectives down to 1e-13 of the value at x0 (one row: the first step lands on 0)
    ident = np.arange(n)
    for rule, gd, mkg in cases:
        mk = {"our": lambda M: M.OurRule(gamma=1 / Lf), "mm": lambda M: M.MalitskyMishchenkoRule(gamma=1 / Lf),
              "fixed": lambda M: M.FixedStepsize(1 / Lf), "plus": lambda M: M.OurRulePlus(gamma=1 / Lf)}[rule]
        logo = []
        xo, ito = O.adaptive_proxgrad(x0, f=O.LinearLeastSquares(A, b), g=mkg(ident), rule=mk(O), tol=1e-9, maxit=60, log=logo)
        # the oracle's own sensitivity to the summation order: the same problem with its columns permuted (three samples)
        gperms = []
        for sd in range(3):
            pm = np.random.default_rng(sd).permutation(n)
            lp = []
            O.adaptive_proxgrad(x0[pm], f=O.LinearLeastSquares(np.asfortranarray(A[:, pm]), b), g=mkg(pm), rule=mk(O), tol=1e-9, maxit=60, log=lp)
            gperms.append(_gam(lp))
        kenv = min(len(logo), min(len(g_) for g_ in gperms))
        env = drift.perm_envelope(_gam(logo, kenv), [g_[:kenv] for g_ in gperms]) if kenv > 0 else np.zeros(0)
        got = {}
        for mode, env_ in (("1", env_on), ("0", env_off)):
            os.environ.update(env_)
            try:
                fc = AdaProx.Counting(f)
                log = []
                x, it = AdaProx.adaptive_proxgrad(x0, f=fc, g=gd, rule=mk(AdaProx), tol=1e-9, maxit=60, log=log)
                got[mode] = (x, it, log, AdaProx.last_solve_info()["matrix_passes"], (fc.eval_count, fc.grad_count))
            finally:
                for k_ in env_:
                    os.environ.pop(k_, None)
        assert got["1"][3] == passes_on and got["0"][3] == 2, ("kernel selection", got["1"][3], got["0"][3])
        for mode in ("1", "0"):
            x, it, log, _, counts = got[mode]
            assert abs(it - ito) <= 1 and counts == (it + 1, it + 1), (rule, mode, it, ito)
            k = min(25, len(log), kenv)
            dd = np.abs(_gam(log, k) / _gam(logo, k) - 1)
            assert np.all(dd <= np.maximum(1e-12, 20 * env[:k])), (rule, mode, float(dd.max()), float(env[:k].max() if k else 0))
            k10 = min(k, 10)
            # (absolute floor relative to the first value: on underdetermined instances the objective runs to 0 through cancellation)
            assert np.allclose([r["objective"] for r in log[:k10]], [r["objective"] for r in logo[:k10]], rtol=1e-10,
                               atol=1e-13 * max(abs(logo[0]["objective"]), fscale)), (rule, mode)
            assert np.allclose([r["norm_res"] for r in log[:k10]], [r["norm_res"] for r in logo[:k10]], rtol=1e-9,
                               atol=1e-12 * gscale), (rule, mode)
            assert np.linalg.norm(x - xo) <= 1e-6 * max(np.linalg.norm(xo), 1e-12), (rule, mode)
    for maxit in (0, 1):
        os.environ.update(env_on)
        try:
            x, it = AdaProx.adaptive_proxgrad(x0, f=f, g=AdaProx.NormL1(0.3), rule=AdaProx.OurRule(gamma=1 / Lf), tol=0.0, maxit=maxit)
            assert AdaProx.last_solve_info()["matrix_passes"] == passes_on
        finally:
            for k_ in env_on:
                os.environ.pop(k_, None)
        xo, ito = O.adaptive_proxgrad(x0, f=O.LinearLeastSquares(A, b), g=O.NormL1(0.3), rule=O.OurRule(gamma=1 / Lf), tol=0.0, maxit=maxit)
        assert it == ito == maxit and np.linalg.norm(x - xo) <= 1e-13 * max(np.linalg.norm(xo), 1e-12)
    f.mat.free()
