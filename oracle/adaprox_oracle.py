"""CPU oracle: a numpy restatement of AdaProx.jl's Float64 algorithms.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the
checker (or as the timed CPU baseline), never as a fallback for the CUDA path.

PARITY UNPINNED at the third-party boundary: the reference is 100 % Julia and
Julia is not installed in the build container, the reference ships no golden
vectors, and the prox bodies live in ProximalCore / ProximalOperators, which are
neither vendored nor version-pinned (Project.toml has no [compat], Manifest.toml
is git-ignored).  The oracle is therefore pinned against first principles
instead (tests/test_oracle_*.py): brute-force argmin for every prox, the Moreau
identity, the closed-form optimum of the Nesterov worst-case function, the
planted-lasso KKT optimum, the reference's own test inequalities on the 2-D toy
(test/runtests.jl:29-50) and its counter equalities (test/runtests.jl:64-89).

Every function cites the reference lines it follows (paths relative to
/root/reference).  Operation order follows the Julia text: temporaries are
formed where Julia forms them, ``norm(v)^2`` is a square root followed by a
square, ``min`` propagates NaN like Julia's, division by zero yields Inf/NaN.
"""
from __future__ import annotations


import numpy as np

F64 = np.float64
_ERR = dict(divide="ignore", invalid="ignore", over="ignore")


class precision:
    """``with precision(np.longdouble): ...`` runs the SAME restatement with every scalar and every temporary in x87
    extended precision (64-bit mantissa, eps = 1.1e-19) instead of Float64: callers pass ``longdouble`` arrays and build
    their rule / prox objects inside the block.  This is not what the reference computes -- it is the yardstick that
    separates "our rounding" from "their rounding" (hard parts (d)): the Float64 oracle's own distance
    from the extended-precision trajectory is the intrinsic rounding drift of the algorithm, and a device trajectory
    that stays inside a small multiple of it is as close to the reference as any Float64 evaluation order can be
    (oracle/drift.py, tests/test_gpu_parity.py)."""

    def __init__(self, dtype):
        self.dtype = dtype

    def __enter__(self):
        global F64
        self._saved = F64
        F64 = self.dtype
        return self

    def __exit__(self, *exc):
        global F64
        F64 = self._saved
        return False


# --------------------------------------------------------------------------
# Julia scalar semantics
# --------------------------------------------------------------------------

def jl_min(*args):
    """Julia ``min``: NaN if any argument is NaN (unlike C fmin / numpy fmin)."""
    out = F64(args[0])
    for a in args[1:]:
        a = F64(a)
        if np.isnan(out) or np.isnan(a):
            out = F64(np.nan)
        else:
            out = a if a < out else out
    return out


def jl_max(*args):
    out = F64(args[0])
    for a in args[1:]:
        a = F64(a)
        if np.isnan(out) or np.isnan(a):
            out = F64(np.nan)
        else:
            out = a if a > out else out
    return out


def jl_sqrt(v):
    """Julia ``sqrt`` throws DomainError on negative reals."""
    v = F64(v)
    if v < 0:
        raise ValueError(f"DomainError: sqrt({v})")
    return F64(np.sqrt(v))


def nan_to_zero(v):
    """src/AdaProx.jl:24 -- only NaN maps to zero; +-Inf passes through."""
    v = F64(v)
    return F64(0.0) if np.isnan(v) else v


def norm(v):
    """LinearAlgebra.norm of a vector (2-norm) / Frobenius norm of a matrix."""
    v = np.asarray(v, dtype=F64)
    return F64(np.sqrt(np.dot(v.ravel(), v.ravel())))


def dot(a, b):
    return F64(np.dot(np.asarray(a, dtype=F64).ravel(), np.asarray(b, dtype=F64).ravel()))


def upper_bound(x, f_x, grad_x, z, gamma):
    """src/AdaProx.jl:26 -- descent-lemma quadratic model."""
    with np.errstate(**_ERR):
        return f_x + dot(grad_x, z - x) + F64(1) / (2 * gamma) * norm(z - x) ** 2


# --------------------------------------------------------------------------
# Counting (src/counting.jl)
# --------------------------------------------------------------------------

_counting_enabled = True


def is_counting_enabled():
    return _counting_enabled


class without_counting:
    """src/counting.jl:8-14.  Julia's do-block form becomes a context manager."""

    def __enter__(self):
        global _counting_enabled
        _counting_enabled = False

    def __exit__(self, *exc):
        global _counting_enabled
        _counting_enabled = True
        return False


class Counting:
    """src/counting.jl:16-33: wrapper with five integer counters."""

    def __init__(self, f):
        self.f = f
        self.eval_count = 0
        self.grad_count = 0
        self.prox_count = 0
        self.mul_count = 0
        self.amul_count = 0

    def __call__(self, *args):          # counting.jl:34
        return self.f(*args)

    # counting.jl:36-51
    def eval_with_pullback(self, x):
        if is_counting_enabled():
            self.eval_count += 1
        f_x, pb = eval_with_pullback(self.f, x)

        def counting_pullback():
            if is_counting_enabled():
                self.grad_count += 1
            return pb()

        return f_x, counting_pullback

    # counting.jl:53-59
    def prox_into(self, x, gamma):
        if is_counting_enabled():
            self.prox_count += 1
        return _prox_into(self.f, x, gamma)

    # counting.jl:25-27 lets convex_conjugate(Counting(h)) go through the
    # generic (Moreau) path; the inner prox call is the one that is counted.
    is_counting = True

    # counting.jl:65-66
    def norm(self):
        return norm(self.f)

    @property
    def T(self):                        # counting.jl:66  adjoint(C)
        return AdjointCounting(self)

    # counting.jl:68-74
    def __matmul__(self, x):
        if is_counting_enabled():
            self.mul_count += 1
        return mul(self.f, x)


class AdjointCounting:
    """src/counting.jl:61-63,76-82."""

    def __init__(self, op):
        self.op = op

    def __matmul__(self, x):
        if is_counting_enabled():
            self.op.amul_count += 1
        return amul(self.op.f, x)


def _count(c, name):
    return getattr(c, name) if isinstance(c, Counting) else None


def grad_count(c):   # counting.jl:84-85
    return _count(c, "grad_count")


def prox_count(c):   # counting.jl:87-88
    return _count(c, "prox_count")


def mul_count(c):    # counting.jl:90-91
    return _count(c, "mul_count")


def amul_count(c):   # counting.jl:93-94
    return _count(c, "amul_count")


def eval_count(c):   # counting.jl:96-97
    return _count(c, "eval_count")


# --------------------------------------------------------------------------
# Linear-operator protocol: ``A * x`` and ``A' * y``  (src/AdaProx.jl:327,329)
# --------------------------------------------------------------------------

def mul(A, x):
    if isinstance(A, Counting):
        return A @ x
    if np.isscalar(A):                   # A = 0 in adaptive_proxgrad (:419)
        return F64(A) * x
    return A @ x                         # ndarray or scipy.sparse


def amul(A, y):
    if isinstance(A, Counting):
        return A.T @ y
    if np.isscalar(A):
        return F64(A) * y
    return A.T @ y


# --------------------------------------------------------------------------
# Gradient-oracle protocol  (src/AdaProx.jl:11-16)
# --------------------------------------------------------------------------

def eval_with_pullback(f, x):
    if not hasattr(f, "eval_with_pullback"):
        raise TypeError(f"eval_with_pullback not defined for type {type(f).__name__}")
    return f.eval_with_pullback(x)


def eval_with_gradient(f, x):
    f_x, pb = eval_with_pullback(f, x)
    return f_x, pb()


# --------------------------------------------------------------------------
# ProximalCore / ProximalOperators semantics (sources are
# external to the reference and unpinned -> checked from first principles)
# --------------------------------------------------------------------------

class Zero:
    """ProximalCore.Zero: f = 0, prox = identity.  Also used as a smooth term
    with a zero gradient (experiments/least_absolute_deviation/runme.jl:18-21)."""

    def __call__(self, x):
        return F64(0.0)

    def prox_into(self, x, gamma):
        return np.array(x, dtype=F64, copy=True), F64(0.0)

    def eval_with_pullback(self, x):
        return self(x), (lambda: np.zeros_like(x))


class IndZero:
    """ProximalCore.IndZero: indicator of {0}; prox = 0."""

    def __call__(self, x):
        return F64(0.0) if not np.any(x) else F64(np.inf)

    def prox_into(self, x, gamma):
        return np.zeros_like(x, dtype=F64), F64(0.0)


class NormL1:
    """ProximalOperators.NormL1(lambda): soft threshold, written branch-wise
    exactly as ``y_i = x_i + (x_i <= -gl ? gl : (x_i >= gl ? -gl : -x_i))``."""

    def __init__(self, lam=1.0):
        self.lam = F64(lam)

    def __call__(self, x):
        return self.lam * F64(np.sum(np.abs(x)))

    def prox_into(self, x, gamma):
        gl = F64(gamma) * self.lam
        y = x + np.where(x <= -gl, gl, np.where(x >= gl, -gl, -x))
        return y, self.lam * F64(np.sum(np.abs(y)))


class NormL2:
    """ProximalOperators.NormL2(lambda): block soft threshold."""

    def __init__(self, lam=1.0):
        self.lam = F64(lam)

    def __call__(self, x):
        return self.lam * norm(x)

    def prox_into(self, x, gamma):
        with np.errstate(**_ERR):
            normx = norm(x)
            scale = jl_max(F64(0.0), F64(1.0) - self.lam * F64(gamma) / normx)
            y = scale * x
        return y, self.lam * scale * normx


class IndBox:
    """ProximalOperators.IndBox(lo, hi) with scalar or per-coordinate bounds."""

    def __init__(self, lo, hi):
        self.lo = lo
        self.hi = hi

    def __call__(self, x):
        return F64(0.0) if np.all((x >= self.lo) & (x <= self.hi)) else F64(np.inf)

    def prox_into(self, x, gamma):
        y = np.where(x < self.lo, self.lo, np.where(x > self.hi, self.hi, x)).astype(F64)
        return y, F64(0.0)


class Translate:
    """ProximalOperators.Translate(f, b): x -> f(x + b)."""

    def __init__(self, f, b):
        self.f = f
        self.b = np.asarray(b, dtype=F64)

    def __call__(self, x):
        return self.f(x + self.b)

    def prox_into(self, x, gamma):
        z = x + self.b
        y, v = _prox_into(self.f, z, gamma)
        y = y - self.b
        return y, v


class ConvexConjugate:
    """ProximalCore.ConvexConjugate: prox through the Moreau identity in the
    order ``u = x ./ gamma; v = prox!(y, f, u, 1/gamma); y .= x .- gamma .* y``."""

    def __init__(self, f):
        self.f = f

    def prox_into(self, x, gamma):
        gamma = F64(gamma)
        with np.errstate(**_ERR):
            u = x / gamma
            y, v = _prox_into(self.f, u, F64(1.0) / gamma)
            v = dot(x, y) - gamma * dot(y, y) - v
            y = x - gamma * y
        return y, v


def convex_conjugate(h):
    """ProximalCore.convex_conjugate (src/AdaProx.jl:325,492,594)."""
    if isinstance(h, Zero):
        return IndZero()
    if isinstance(h, IndZero):
        return Zero()
    if isinstance(h, ConvexConjugate):      # the biconjugate of a closed convex function is the function itself
        return h.f
    return ConvexConjugate(h)


def _prox_into(f, x, gamma):
    return f.prox_into(x, gamma)


def prox(f, x, gamma=1.0):
    """ProximalCore.prox(f, x, gamma) -> (y, f(y))."""
    return _prox_into(f, np.asarray(x, dtype=F64), F64(gamma))


# --------------------------------------------------------------------------
# Smooth-term oracles defined by the experiment scripts and the test file
# --------------------------------------------------------------------------

class LinearLeastSquares:
    """experiments/lasso/runme.jl:16-27."""

    def __init__(self, A, b):
        self.A = A
        self.b = b

    def eval_with_pullback(self, w):
        res = self.A @ w - self.b
        return F64(0.5) * norm(res) ** 2, (lambda: self.A.T @ res)

    def __call__(self, w):
        return self.eval_with_pullback(w)[0]


class LogisticLoss:
    """experiments/sparse_logreg/runme.jl:18-39 (intercept is w[end]; the
    naive ``1 + exp(-logits)`` form is kept on purpose)."""

    def __init__(self, X, y):
        self.X = X
        self.y = np.asarray(y, dtype=F64)

    def eval_with_pullback(self, w):
        with np.errstate(**_ERR):
            logits = self.X @ w[:-1] + w[-1]
            u = 1 + np.exp(-logits)

            def pullback():
                probs = 1 / u
                N = self.y.shape[0]
                grad = np.zeros_like(w)
                grad[:-1] = (self.X.T @ (probs - self.y)) / N
                grad[-1] = np.mean(probs - self.y)
                return grad

            return -F64(np.mean((self.y - 1) * logits - np.log(u))), pullback

    def __call__(self, w):
        return self.eval_with_pullback(w)[0]


class Quadratic:
    """experiments/dual_svm/runme.jl:19-28."""

    def __init__(self, Q, q):
        self.Q = Q
        self.q = q

    def eval_with_pullback(self, x):
        temp = self.Q @ x
        return F64(0.5) * dot(x, temp) + dot(x, self.q), (lambda: temp + self.q)

    def __call__(self, x):
        return self.eval_with_pullback(x)[0]


class Cubic:
    """experiments/cubic_sparse_logreg/runme.jl:20-32."""

    def __init__(self, Q, q, c):
        self.Q = Q
        self.q = q
        self.c = F64(c)

    def eval_with_pullback(self, x):
        grad = self.Q @ x + self.q + (norm(x) * self.c / 2) * x
        return (dot(x, grad) + dot(self.q, x)) / 2 - norm(x) ** 3 * self.c / 12, (lambda: grad)

    def __call__(self, x):
        return self.eval_with_pullback(x)[0]


def logistic_loss_grad_Hessian(X, y, w):
    """experiments/cubic_sparse_logreg/runme.jl:34-45 (setup of the Cubic oracle): returns (H, g)."""
    y = np.asarray(y, dtype=F64)
    with np.errstate(**_ERR):
        probs = 1 / (1 + np.exp(-(X @ w[:-1] + w[-1])))           # sigm.(X * w[1:end-1] .+ w[end])
        N = y.shape[0]
        g = np.append((X.T @ (probs - y)) / N, np.mean(probs - y))
        sb = probs * (1 - probs) / N
        Xd = X.toarray() if hasattr(X, "toarray") else np.asarray(X)
        XtR = Xd.T * sb                                            # X' * R, R = diagm(sb)
        XR = XtR @ np.ones((N, 1))
        H = np.vstack([np.hstack([XtR @ Xd, XR]), np.hstack([XR.T, [[np.sum(sb)]]])])
    return H, g


class WorstQuadratic:
    """experiments/nesterov_worst_case/runme.jl:14-40 (1-based k -> x[k-1])."""

    def __init__(self, k, L):
        self.k = int(k)
        self.L = F64(L)

    def eval_with_pullback(self, x):
        k, L = self.k, self.L
        s = x[0] ** 2 + x[k - 1] ** 2
        for i in range(k - 1):
            s += (x[i] - x[i + 1]) ** 2

        def pullback():
            grad = np.zeros_like(x)
            grad[0] = (L / 4) * (2 * x[0] - x[1] - 1)
            for i in range(1, k - 1):
                grad[i] = (L / 4) * (2 * x[i] - x[i - 1] - x[i + 1])
            grad[k - 1] = (L / 4) * (2 * x[k - 1] - x[k - 2])
            grad[k:] = 0
            return grad

        return (L / 4) * (s / 2 - x[0]), pullback

    def __call__(self, x):
        return self.eval_with_pullback(x)[0]


class Simple2DObjective:
    """test/runtests.jl:6-13."""

    def eval_with_pullback(self, x):
        def pullback():
            return np.array([2 * np.log(1 + x[0] ** 2) * 2 * x[0] / (1 + x[0] ** 2), 20 * x[1]], dtype=F64)

        return F64(np.log(1 + x[0] ** 2) ** 2 + 10 * x[1] ** 2), pullback

    def __call__(self, x):
        return self.eval_with_pullback(x)[0]


class Simple2DBox:
    """test/runtests.jl:15-23."""

    def __call__(self, x):
        return F64(0.0) if abs(x[0]) <= 2.9 else F64(np.inf)

    def prox_into(self, x, gamma):
        y = np.array(x, dtype=F64, copy=True)
        y[0] = min(max(x[0], -2.9), 2.9)
        return y, F64(0.0)


# --------------------------------------------------------------------------
# Record emission (the ``@logmsg Record`` lines).  ``log`` is None (default
# logger: the keyword expressions are never evaluated) or a list that receives
# one dict per iteration, evaluated inside ``without_counting``.
# --------------------------------------------------------------------------

def _emit(log, **kw):
    if log is not None:
        log.append(kw)


# --------------------------------------------------------------------------
# Backtracking proximal gradient / Nesterov  (src/AdaProx.jl:34-84)
# --------------------------------------------------------------------------

def backtrack_stepsize(gamma, f, g, x, f_x, grad_x, shrink=0.5, warn=None):
    """src/AdaProx.jl:34-48."""
    gamma = F64(gamma)
    z, g_z = prox(g, x - gamma * grad_x, gamma)
    ub_z = upper_bound(x, f_x, grad_x, z, gamma)
    f_z, pb = eval_with_pullback(f, z)
    while f_z > ub_z:
        gamma = gamma * shrink
        if gamma < 1e-12:                # :40-42 logs an error and keeps going
            if warn is not None:
                warn.append(float(gamma))
            if gamma < 1e-300:           # the Julia loop would spin forever
                raise FloatingPointError("step size underflow in backtrack_stepsize")
        z, g_z = prox(g, x - gamma * grad_x, gamma)
        ub_z = upper_bound(x, f_x, grad_x, z, gamma)
        f_z, pb = eval_with_pullback(f, z)
    return gamma, z, f_z, g_z, pb


def backtracking_proxgrad(x0, *, f, g, gamma0, xi=1.0, shrink=0.5, tol=1e-5, maxit=100_000,
                          name="Backtracking PG", log=None):
    """src/AdaProx.jl:50-64."""
    x, z, gamma = x0, x0, F64(gamma0)
    f_x, grad_x = eval_with_gradient(f, x)
    for it in range(1, int(maxit) + 1):
        gamma, z, f_z, g_z, pb = backtrack_stepsize(xi * gamma, f, g, x, f_x, grad_x, shrink)
        norm_res = norm(z - x) / gamma
        _emit(log, method=name, it=it, gamma=gamma, norm_res=norm_res, objective=f_z + g_z,
              grad_f_evals=grad_count(f), prox_g_evals=prox_count(g), f_evals=eval_count(f))
        if norm_res <= tol:
            return z, it
        x, f_x = z, f_z
        grad_x = pb()
    return z, maxit


def backtracking_nesterov(x0, *, f, g, gamma0, shrink=0.5, tol=1e-5, maxit=100_000,
                          name="Backtracking Nesterov", log=None):
    """src/AdaProx.jl:66-84."""
    x, z, gamma = x0, x0, F64(gamma0)
    theta = F64(1.0)
    f_x, grad_x = eval_with_gradient(f, x)
    for it in range(1, int(maxit) + 1):
        z_prev = z
        gamma, z, f_z, g_z, _ = backtrack_stepsize(gamma, f, g, x, f_x, grad_x, shrink)
        norm_res = norm(z - x) / gamma
        _emit(log, method=name, it=it, gamma=gamma, norm_res=norm_res, objective=f_z + g_z,
              grad_f_evals=grad_count(f), prox_g_evals=prox_count(g), f_evals=eval_count(f))
        if norm_res <= tol:
            return z, it
        theta_prev = theta
        theta = (1 + jl_sqrt(1 + 4 * theta_prev ** 2)) / 2
        x = z + (theta_prev - 1) / theta * (z - z_prev)
        f_x, grad_x = eval_with_gradient(f, x)
    return z, maxit


# --------------------------------------------------------------------------
# Fixed-step accelerated proximal gradient  (src/AdaProx.jl:91-142)
# --------------------------------------------------------------------------

def fixed_nesterov(x0, *, f, g, Lf=None, muf=0, mug=0, gamma=None, theta=None, tol=1e-5,
                   maxit=100_000, name="Fixed Nesterov", log=None):
    assert (gamma is None) != (Lf is None)                       # :104
    if gamma is None:
        gamma = F64(1) / F64(Lf)
    gamma = F64(gamma)
    mu = F64(muf) + F64(mug)
    q = gamma * mu / (1 + gamma * mug)
    assert q < 1                                                  # :110
    if theta is None:
        theta = F64(1) / jl_sqrt(q) if q > 0 else F64(0)
    theta = F64(theta)
    with np.errstate(**_ERR):
        assert 0 <= theta <= F64(1) / jl_sqrt(q)                 # :118 (1/sqrt(0) = Inf)
    x, x_prev = x0, x0
    for it in range(1, int(maxit) + 1):
        theta_prev = theta
        if mu == 0:
            theta = (1 + jl_sqrt(1 + 4 * theta_prev ** 2)) / 2
            beta = (theta_prev - 1) / theta
        else:
            theta = (1 - q * theta_prev ** 2 + jl_sqrt((1 - q * theta_prev ** 2) ** 2 + 4 * theta_prev ** 2)) / 2
            beta = (theta_prev - 1) * (1 + gamma * mug - theta * gamma * mu) / theta / (1 - gamma * muf)
        z = x + beta * (x - x_prev)
        _, grad_z = eval_with_gradient(f, z)
        x_prev = x
        x, g_x = prox(g, z - gamma * grad_z, gamma)
        norm_res = norm(x - z) / gamma
        if log is not None:
            with without_counting():
                _emit(log, method=name, it=it, gamma=gamma, norm_res=norm_res, objective=f(x) + g_x,
                      grad_f_evals=grad_count(f), prox_g_evals=prox_count(g), f_evals=eval_count(f))
        if norm_res <= tol:
            return x, it
    return x, maxit


# --------------------------------------------------------------------------
# aGRAAL  (src/AdaProx.jl:150-192)
# --------------------------------------------------------------------------

def agraal(x1, *, f, g, x0=None, gamma0=None, gamma_max=1e6, phi=1.5, tol=1e-5, maxit=100_000,
           name="aGRAAL", log=None, rng=None):
    if x0 is None:                                                # :162-164 (randn)
        rng = np.random.default_rng(0) if rng is None else rng
        x0 = x1 + rng.standard_normal(x1.shape)
    x, x_prev, x_bar = x1, x0, x1
    _, grad_x = eval_with_gradient(f, x)
    _, grad_x_prev = eval_with_gradient(f, x_prev)
    with np.errstate(**_ERR):
        if gamma0 is None:
            gamma0 = norm(x - x_prev) / norm(grad_x - grad_x_prev)
        gamma = F64(gamma0)
        phi = F64(phi)
        rho = 1 / phi + 1 / phi ** 2
        theta = F64(1.0)
        for it in range(1, int(maxit) + 1):
            C = norm(x - x_prev) ** 2 / norm(grad_x - grad_x_prev) ** 2
            gamma_prev = gamma
            gamma = jl_min(rho * gamma_prev, phi * theta * C / (4 * gamma_prev), gamma_max)
            theta = phi * gamma / gamma_prev
            x_bar = ((phi - 1) * x + x_bar) / phi
            x_prev, grad_x_prev = x, grad_x
            x, g_x = prox(g, x_bar - gamma * grad_x_prev, gamma)
            norm_res = norm(x - x_prev) / gamma
            if log is not None:
                with without_counting():
                    _emit(log, method=name, it=it, gamma=gamma, norm_res=norm_res, objective=f(x) + g_x,
                          grad_f_evals=grad_count(f), prox_g_evals=prox_count(g), f_evals=eval_count(f))
            if norm_res <= tol:
                return x, it
            _, grad_x = eval_with_gradient(f, x)
    return x, maxit


# --------------------------------------------------------------------------
# Stepsize rules  (src/AdaProx.jl:208-308)
# --------------------------------------------------------------------------

class FixedStepsize:
    """src/AdaProx.jl:208-215."""

    def __init__(self, gamma, t=1.0):
        self.gamma = F64(gamma)
        self.t = F64(t)

    def stepsize(self, *args):
        return (self.gamma, self.gamma * self.t ** 2), None


class MalitskyMishchenkoRule:
    """src/AdaProx.jl:217-230."""

    def __init__(self, gamma, t=1.0):
        self.gamma = F64(gamma)
        self.t = F64(t)

    def stepsize(self, state=None, x1=None, grad_x1=None, x0=None, grad_x0=None):
        if state is None:                                         # :222-224
            return (self.gamma, self.gamma * self.t ** 2), (self.gamma, F64(np.inf))
        gamma_prev, rho = state                                   # :226-230
        with np.errstate(**_ERR):
            L = norm(grad_x1 - grad_x0) / norm(x1 - x0)
            gamma = jl_min(jl_sqrt(1 + rho) * gamma_prev, F64(1) / (2 * L))
            return (gamma, gamma * self.t ** 2), (gamma, gamma / gamma_prev)


class OurRule:
    """src/AdaProx.jl:232-273."""

    def __init__(self, gamma=0, t=1, norm_A=0, delta=0, Theta=1.2):
        if gamma > 0:                                             # :241-247
            _gamma = F64(gamma)
        elif norm_A > 0:
            _gamma = F64(1) / (2 * F64(Theta) * F64(t) * F64(norm_A))
        else:
            raise ValueError("you must provide gamma > 0 if norm_A = 0")
        self.gamma = _gamma
        self.t = F64(t)
        self.norm_A = F64(norm_A)
        self.delta = F64(delta)
        self.Theta = F64(Theta)

    def stepsize(self, state=None, x1=None, grad_x1=None, x0=None, grad_x0=None):
        if state is None:                                         # :252-256
            gamma = self.gamma
            sigma = self.gamma * self.t ** 2
            return (gamma, sigma), (gamma, gamma)
        gamma1, gamma0 = state                                    # :258-273
        with np.errstate(**_ERR):
            xi = self.t ** 2 * gamma1 ** 2 * self.norm_A ** 2
            C = nan_to_zero(norm(grad_x1 - grad_x0) ** 2 / dot(grad_x1 - grad_x0, x1 - x0))
            L = nan_to_zero(dot(grad_x1 - grad_x0, x1 - x0) / norm(x1 - x0) ** 2)
            D = gamma1 * L * (gamma1 * C - 1)
            d1 = 1 + self.delta
            gamma = jl_min(
                gamma1 * jl_sqrt(1 + gamma1 / gamma0),
                F64(1) / (2 * self.Theta * self.t * self.norm_A),
                (
                    gamma1 * jl_sqrt(1 - 4 * xi * d1 ** 2)
                    / jl_sqrt(2 * d1 * (D + jl_sqrt(D ** 2 + xi * (1 - 4 * xi * d1 ** 2))))
                ),
            )
            sigma = gamma * self.t ** 2
        return (gamma, sigma), (gamma, gamma1)


class OurRulePlus:
    """src/AdaProx.jl:277-308."""

    def __init__(self, gamma=0, nu=1, xi=1, r=0.5):
        if not gamma > 0:
            raise ValueError("you must provide gamma > 0")
        self.gamma = F64(gamma)
        self.xi = F64(xi)
        self.nu = F64(nu)
        self.r = F64(r)

    def stepsize(self, state=None, x1=None, grad_x1=None, x0=None, grad_x0=None):
        if state is None:                                         # :294-297
            return (self.gamma, self.gamma), (self.gamma, self.gamma)
        gamma1, gamma0 = state                                    # :299-308
        r, nu, xi = self.r, self.nu, self.xi
        with np.errstate(**_ERR):
            C = nan_to_zero(norm(grad_x1 - grad_x0) ** 2 / dot(grad_x1 - grad_x0, x1 - x0))
            L = nan_to_zero(dot(grad_x1 - grad_x0, x1 - x0) / norm(x1 - x0) ** 2)
            D = nan_to_zero(1 - 2 * r + gamma1 * L * (gamma1 * C + 2 * (r - 1)))
            gamma = gamma1 * jl_min(
                jl_sqrt(1 / (r * (nu + xi)) + gamma1 / gamma0),
                jl_sqrt((nu * (1 + xi) - 1) / (nu * (nu + xi))) / jl_sqrt(jl_max(D, 0)),
            )
        return (gamma, gamma), (gamma, gamma1)


def stepsize(rule, *args):
    return rule.stepsize(*args)


# --------------------------------------------------------------------------
# The generic adaptive primal-dual loop  (src/AdaProx.jl:312-364) -- AdaPDM
# --------------------------------------------------------------------------

def adaptive_primal_dual(x, y, *, f, g, h, A, rule, tol=1e-5, maxit=10_000, name="AdaPDM",
                         log=None, trace=None):
    """``trace`` (test hook, not in the reference): list receiving the complete
    loop state at the top of every iteration, for teacher-forced comparisons."""
    (gamma, sigma), state = stepsize(rule)                       # :324
    h_conj = convex_conjugate(h)                                  # :325

    with np.errstate(**_ERR):
        A_x = mul(A, x)                                           # :327
        _, grad_x = eval_with_gradient(f, x)                      # :328
        At_y = amul(A, y)                                         # :329
        v = x - gamma * (grad_x + At_y)                           # :330
        x_prev, A_x_prev, grad_x_prev = x, A_x, grad_x            # :331
        x, _ = prox(g, v, gamma)                                  # :332

        for it in range(1, int(maxit) + 1):                       # :334
            A_x = mul(A, x)                                       # :335
            f_x, grad_x = eval_with_gradient(f, x)                # :336

            primal_res = (v - x) / gamma + grad_x + At_y          # :338

            gamma_prev = gamma                                    # :340
            (gamma, sigma), state = stepsize(rule, state, x, grad_x, x_prev, grad_x_prev)  # :341
            rho = gamma / gamma_prev                              # :342

            w = y + sigma * ((1 + rho) * A_x - rho * A_x_prev)    # :344
            y, _ = prox(h_conj, w, sigma)                         # :345

            dual_res = (w - y) / sigma - A_x                      # :347
            norm_res = jl_sqrt(norm(primal_res) ** 2 + norm(dual_res) ** 2)   # :348

            if log is not None:                                   # :350-352
                with without_counting():
                    _emit(log, method=name, it=it, gamma=gamma, sigma=sigma, norm_res=norm_res,
                          objective=f_x + g(x) + h(A_x), grad_f_evals=grad_count(f),
                          prox_g_evals=prox_count(g), prox_h_evals=prox_count(h),
                          A_evals=mul_count(A), At_evals=amul_count(A), f_evals=eval_count(f))
            if trace is not None:
                trace.append(dict(it=it, x=x.copy(), grad_x=np.array(grad_x, copy=True), f_x=f_x,
                                  gamma=gamma, sigma=sigma, norm_res=norm_res, y=np.array(y, copy=True)))

            if norm_res <= tol:                                   # :354-356
                return x, y, it

            At_y = amul(A, y)                                     # :358
            v = x - gamma * (grad_x + At_y)                       # :359
            x_prev, A_x_prev, grad_x_prev = x, A_x, grad_x        # :360
            x, _ = prox(g, v, gamma)                              # :361
    return x, y, maxit


def condat_vu(x, y, *, f, g, h, A, Lf, gamma=None, sigma=None, norm_A=None, tol=1e-5,
              maxit=10_000, name="Condat-Vu", log=None):
    """src/AdaProx.jl:367-416."""
    if gamma is None and sigma is None:                           # :398-412
        Lf = F64(Lf)
        par = F64(5)
        par2 = F64(100)
        if norm_A is None:
            norm_A = A.norm() if isinstance(A, Counting) else norm(A)
        norm_A = F64(norm_A)
        with np.errstate(**_ERR):
            if norm_A > par * Lf:
                alpha = F64(1)
            else:
                alpha = par2 * norm_A / Lf
            gamma = F64(1) / (Lf / 2 + norm_A / alpha)
            sigma = F64(0.99) / (norm_A * alpha)
    assert gamma is not None and sigma is not None               # :413
    rule = FixedStepsize(gamma, jl_sqrt(F64(sigma) / F64(gamma)))  # :414
    return adaptive_primal_dual(x, y, f=f, g=g, h=h, A=A, rule=rule, tol=tol, maxit=maxit,
                                name=name, log=log)


def adaptive_proxgrad(x, *, f, g, rule, tol=1e-5, maxit=100_000, name="AdaPGM", log=None, trace=None):
    """src/AdaProx.jl:418-421."""
    x, _, numit = adaptive_primal_dual(x, np.zeros_like(x), f=f, g=g, h=Zero(), A=0, rule=rule,
                                       tol=tol, maxit=maxit, name=name, log=log, trace=trace)
    return x, numit


def auto_adaptive_proxgrad(x, *, f, g, gamma=None, tol=1e-5, maxit=100_000, name="AutoAdaPGM", log=None):
    """src/AdaProx.jl:423-455.  The ``gamma === nothing`` branch of the
    reference cannot run (:431 calls ``prox`` without ``g``); it raises here."""
    _, grad_x = eval_with_gradient(f, x)
    if norm(grad_x) <= tol:
        return x, 0
    if gamma is None:
        raise TypeError("reference :431 calls prox(x, gamma) without g -> MethodError")
    assert gamma > 0
    gamma = F64(gamma)
    with np.errstate(**_ERR):
        x_prev, grad_x_prev, gamma_prev = x, grad_x, gamma
        x, _ = prox(g, x - gamma * grad_x, gamma)
        _, grad_x = eval_with_gradient(f, x)
        L = dot(grad_x - grad_x_prev, x - x_prev) / norm(x - x_prev) ** 2
        gamma = jl_sqrt(2) * gamma if L == 0 else F64(1) / L
        if gamma_prev / gamma > 1e5:
            x, _ = prox(g, x_prev - gamma * grad_x_prev, gamma)
            _, grad_x = eval_with_gradient(f, x)
            L = dot(grad_x - grad_x_prev, x - x_prev) / norm(x - x_prev) ** 2
            gamma = jl_sqrt(2) * gamma if L == 0 else F64(1) / L
    rule = OurRule(gamma=gamma, t=1, norm_A=0, delta=0, Theta=1.2)
    return adaptive_proxgrad(x_prev, f=f, g=g, rule=rule, tol=tol, maxit=maxit, name=name, log=log)


def adaptive_proxgrad_path(X0, *, f, lambdas, rule_of, tol=1e-5, maxit=100_000, history=0):
    """Checker for the batched multi-lambda path (include/adaprox.h: adaprox_solve_lambda_path): the reference has no
    batched entry point, so this is literally one ``adaptive_proxgrad`` call (src/AdaProx.jl:418-421) per lambda.
    ``rule_of(j)`` returns the rule of column j.  Returns (X, its, gamma_hist, res_hist, obj_hist)."""
    lambdas = np.asarray(lambdas, dtype=float)
    n, Lc = X0.shape
    X = np.empty((n, Lc))
    its = np.zeros(Lc, dtype=np.int64)
    H = int(min(history, maxit)) if history else 0
    gh = np.full((H, Lc), np.nan); rh = np.full((H, Lc), np.nan); oh = np.full((H, Lc), np.nan)
    for j in range(Lc):
        log = []
        x, it = adaptive_proxgrad(X0[:, j].copy(), f=f, g=NormL1(float(lambdas[j])), rule=rule_of(j), tol=tol, maxit=maxit, log=log)
        X[:, j] = x
        its[j] = it
        for r in log[:H]:
            gh[r["it"] - 1, j] = r["gamma"]; rh[r["it"] - 1, j] = r["norm_res"]; oh[r["it"] - 1, j] = r["objective"]
    return X, its, gh, rh, oh


def fixed_proxgrad(x, *, f, g, gamma, tol=1e-5, maxit=100_000, name="Fixed stepsize PGM", log=None):
    """src/AdaProx.jl:457-459."""
    return adaptive_proxgrad(x, f=f, g=g, rule=FixedStepsize(gamma, 1.0), tol=tol, maxit=maxit,
                             name=name, log=log)


# --------------------------------------------------------------------------
# AdaPDM+ : linesearch on ||A||  (src/AdaProx.jl:463-550)
# --------------------------------------------------------------------------

def adaptive_linesearch_primal_dual(x, y, *, f, g, h, A, gamma=None, eta=1.0, t=1.0, delta=1e-8,
                                    Theta=1.2, r=2, R=0.95, tol=1e-5, maxit=10_000, name="AdaPDM+",
                                    log=None, trials=None):
    """``trials`` (test hook): list receiving the number of linesearch trials
    of every iteration."""
    eta, t, delta, Theta, r, R = F64(eta), F64(t), F64(delta), F64(Theta), F64(r), F64(R)
    assert eta > 0, "eta must be positive"                       # :481
    assert Theta > (delta + 1), "must be Theta > (delta + 1)"    # :482
    if gamma is None:
        gamma = F64(1) / (2 * Theta * t * eta)                    # :485
    gamma = F64(gamma)
    assert gamma <= F64(1) / (2 * Theta * t * eta), "gamma is too large"   # :488

    delta1 = 1 + delta
    gamma_prev = gamma
    h_conj = convex_conjugate(h)

    with np.errstate(**_ERR):
        A_x = mul(A, x)                                           # :494
        _, grad_x = eval_with_gradient(f, x)
        At_y = amul(A, y)
        v = x - gamma * (grad_x + At_y)
        x_prev, A_x_prev, grad_x_prev = x, A_x, grad_x
        x, _ = prox(g, v, gamma)                                  # :499

        for it in range(1, int(maxit) + 1):
            A_x = mul(A, x)                                       # :502
            f_x, grad_x = eval_with_gradient(f, x)                # :503

            primal_res = (v - x) / gamma + grad_x + At_y          # :505

            C = nan_to_zero(norm(grad_x - grad_x_prev) ** 2 / dot(grad_x - grad_x_prev, x - x_prev))  # :507
            L = nan_to_zero(dot(grad_x - grad_x_prev, x - x_prev) / norm(x - x_prev) ** 2)            # :508
            Delta = gamma * L * (gamma * C - 1)                   # :509
            xi_bar = t ** 2 * gamma ** 2 * eta ** 2 * delta1 ** 2  # :510
            m4xim1 = 1 - 4 * xi_bar                               # :511

            eta = R * eta                                         # :513
            w = y
            sigma = t ** 2 * gamma
            ntrial = 0
            while True:                                           # :516-533
                ntrial += 1
                gamma_next = jl_min(
                    gamma * jl_sqrt(1 + gamma / gamma_prev),
                    F64(1) / (2 * Theta * t * eta),
                    gamma * jl_sqrt(m4xim1 / (2 * delta1 * (Delta + jl_sqrt(Delta ** 2 + m4xim1 * (t * eta * gamma) ** 2)))),
                )
                rho = gamma_next / gamma
                sigma = t ** 2 * gamma_next
                w = y + sigma * ((1 + rho) * A_x - rho * A_x_prev)
                y_next, _ = prox(h_conj, w, sigma)
                At_y_next = amul(A, y_next)
                if eta >= norm(At_y_next - At_y) / norm(y_next - y):
                    gamma, gamma_prev = gamma_next, gamma
                    y, At_y = y_next, At_y_next
                    break
                eta = eta * r
            if trials is not None:
                trials.append(ntrial)

            dual_res = (w - y) / sigma - A_x                      # :535
            norm_res = jl_sqrt(norm(primal_res) ** 2 + norm(dual_res) ** 2)

            if log is not None:                                   # :538-540
                with without_counting():
                    _emit(log, method=name, it=it, gamma=gamma, sigma=sigma, norm_res=norm_res,
                          objective=f_x + g(x) + h(A_x), grad_f_evals=grad_count(f),
                          prox_g_evals=prox_count(g), prox_h_evals=prox_count(h),
                          A_evals=mul_count(A), At_evals=amul_count(A), f_evals=eval_count(f))
            if norm_res <= tol:
                return x, y, it

            v = x - gamma * (grad_x + At_y)                       # :545
            x_prev, A_x_prev, grad_x_prev = x, A_x, grad_x
            x, _ = prox(g, v, gamma)
    return x, y, maxit


# --------------------------------------------------------------------------
# Malitsky-Pock primal-dual linesearch  (src/AdaProx.jl:555-629)
# --------------------------------------------------------------------------

def backtrack_stepsize_MP(sigma, sigma_prev, t, x_prev, y, y_prev, grad_x_prev, A_x_prev, At_y,
                          At_y_prev, f, g, A, f_x_prev):
    """src/AdaProx.jl:555-579."""
    with np.errstate(**_ERR):
        theta = sigma / sigma_prev
        gamma = t ** 2 * sigma
        At_ybar = (1 + theta) * At_y - theta * At_y_prev
        v = x_prev - gamma * (At_ybar + grad_x_prev)
        x, _ = prox(g, v, gamma)
        A_x = mul(A, x)
        f_x, pb = eval_with_pullback(f, x)
        lhs = gamma * sigma * norm(A_x - A_x_prev) ** 2 + 2 * gamma * (f_x - f_x_prev - dot(grad_x_prev, x - x_prev))
        while lhs > 0.95 * norm(x - x_prev) ** 2:
            sigma = sigma / 2
            if sigma < 1e-300:
                raise FloatingPointError("step size underflow in backtrack_stepsize_MP")
            theta = sigma / sigma_prev
            gamma = t ** 2 * sigma
            At_ybar = (1 + theta) * At_y - theta * At_y_prev
            v = x_prev - gamma * (At_ybar + grad_x_prev)
            x, _ = prox(g, v, gamma)
            A_x = mul(A, x)
            f_x, pb = eval_with_pullback(f, x)
            lhs = gamma * sigma * norm(A_x - A_x_prev) ** 2 + 2 * gamma * (f_x - f_x_prev - dot(grad_x_prev, x - x_prev))
    return sigma, gamma, x, v, A_x, f_x, pb


def malitsky_pock(x, y, *, f, g, h, A, sigma, t=1.0, tol=1e-5, maxit=10_000, name="MP-ls", log=None):
    """src/AdaProx.jl:581-629."""
    sigma, t = F64(sigma), F64(t)
    h_conj = convex_conjugate(h)
    theta = F64(1.0)
    y_prev = y
    with np.errstate(**_ERR):
        A_x = mul(A, x)
        At_y = amul(A, y)
        for it in range(1, int(maxit) + 1):
            At_y_prev = At_y
            w = y + sigma * A_x
            y, _ = prox(h_conj, w, sigma)
            At_y = amul(A, y)

            sigma_prev = sigma
            sigma = sigma * jl_sqrt(1 + theta)

            f_x_prev, grad_x_prev = eval_with_gradient(f, x)
            x_prev, A_x_prev = x, A_x
            sigma, gamma, x, v, A_x, f_x, pb = backtrack_stepsize_MP(
                sigma, sigma_prev, t, x_prev, y, y_prev, grad_x_prev, A_x_prev, At_y, At_y_prev, f, g, A, f_x_prev)
            grad_x = pb()

            y_prev = y

            primal_res = (v - x) / gamma + grad_x + At_y
            dual_res = (w - y) / sigma_prev - A_x
            norm_res = jl_sqrt(norm(primal_res) ** 2 + norm(dual_res) ** 2)

            if log is not None:
                with without_counting():
                    _emit(log, method=name, it=it, gamma=gamma, sigma=sigma, norm_res=norm_res,
                          objective=f_x + g(x) + h(A_x), grad_f_evals=grad_count(f),
                          prox_g_evals=prox_count(g), prox_h_evals=prox_count(h),
                          A_evals=mul_count(A), At_evals=amul_count(A), f_evals=eval_count(f))
            if norm_res <= tol:
                return x, y, it
    return x, y, maxit
