/* adaprox_ref.c -- TEST INFRASTRUCTURE: a second, independent CPU restatement of the reference's generic loop in plain C.
 *
 * Written directly from the Julia source (not from oracle/adaprox_oracle.py) so that the numpy oracle -- the checker of every
 * GPU parity test -- is itself cross-checked by an implementation that shares no code, no BLAS and no summation order with it
 * (tests/test_oracle_c_restatement.py).  Only tests/ may load this library; the product never does.
 *
 * Follows, line by line:
 *   adaptive_primal_dual            src/AdaProx.jl:312-364   (AdaPGM :418-421 = the same loop with h = Zero(), A = 0, y = zero(x))
 *   FixedStepsize / MalitskyMishchenkoRule / OurRule          src/AdaProx.jl:208-273
 *   nan_to_zero                                               src/AdaProx.jl:24
 *   LinearLeastSquares   experiments/lasso/runme.jl:21-25     Quadratic   experiments/dual_svm/runme.jl:24-28
 *   LogisticLoss         experiments/sparse_logreg/runme.jl:23-37          Zero  experiments/least_absolute_deviation/runme.jl:18-21
 *   prox bodies: SURVEY.md Appendix A (ProximalCore / ProximalOperators: NormL1, NormL2, IndBox, Zero, IndZero, Translate,
 *   ConvexConjugate through Moreau in ProximalCore's operation order).  Parity unpinned at that boundary (no Julia here).
 *
 * Plain sequential loops, IEEE doubles, no FMA contraction (-ffp-contract=off), matrices column-major like Julia's.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

enum { F_ZERO = 0, F_LEAST_SQUARES = 1, F_LOGISTIC = 2, F_QUADRATIC = 3 };
enum { P_ZERO = 0, P_IND_ZERO = 1, P_NORM_L1 = 2, P_NORM_L2 = 3, P_IND_BOX = 4 };
enum { RULE_FIXED = 0, RULE_MM = 1, RULE_OUR = 2 };

typedef struct {
  int kind;
  double lambda, lo, hi;
  const double* shift;          /* Translate(f, shift): x -> f(x + shift); NULL = none */
} ref_prox;

typedef struct {
  int f_kind;
  const double* F;              /* column-major fm x fn */
  long fm, fn;
  const double* fvec;           /* b / y / q */
  ref_prox g, h;
  const double* A;              /* column-major am x n; NULL = the scalar `A = 0` of adaptive_proxgrad */
  long am, n;
  int rule;
  double gamma, t, norm_A, delta, Theta;
  double tol;
  long maxit;
} ref_problem;

/* ---- Julia scalar semantics ------------------------------------------------------------------------------------ */
static double jl_min(double a, double b) { return (isnan(a) || isnan(b)) ? NAN : (a < b ? a : b); }   /* min propagates NaN */
static double nan_to_zero(double v) { return isnan(v) ? 0.0 : v; }                                    /* :24 */
static double dot(const double* a, const double* b, long n) { double s = 0; for (long i = 0; i < n; ++i) s += a[i] * b[i]; return s; }
static double norm2(const double* a, long n) { return sqrt(dot(a, a, n)); }
static double sq(double v) { return v * v; }

/* ---- dense column-major products --------------------------------------------------------------------------------- */
static void mul(const double* M, long m, long n, const double* x, double* out) {          /* out = M * x */
  for (long i = 0; i < m; ++i) out[i] = 0.0;
  for (long j = 0; j < n; ++j) { const double xj = x[j]; const double* c = M + j * m; for (long i = 0; i < m; ++i) out[i] += c[i] * xj; }
}
static void amul(const double* M, long m, long n, const double* y, double* out) {         /* out = M' * y */
  for (long j = 0; j < n; ++j) out[j] = dot(M + j * m, y, m);
}

/* ---- smooth terms: value and gradient (eval_with_gradient, src/AdaProx.jl:13-16) --------------------------------- */
static double eval_f(const ref_problem* p, const double* x, double* grad, double* tmp /* >= max(fm, n) */) {
  const long n = p->n;
  switch (p->f_kind) {
    case F_LEAST_SQUARES: {                                   /* res = A w - b ; 0.5 norm(res)^2 ; A' res */
      mul(p->F, p->fm, p->fn, x, tmp);
      for (long i = 0; i < p->fm; ++i) tmp[i] -= p->fvec[i];
      const double v = 0.5 * sq(norm2(tmp, p->fm));
      amul(p->F, p->fm, p->fn, tmp, grad);
      return v;
    }
    case F_LOGISTIC: {                                        /* logits = X w[1:end-1] .+ w[end] ; u = 1 + exp(-logits) */
      const long m = p->fm, d = p->fn;                        /* n = d + 1 */
      mul(p->F, m, d, x, tmp);
      double val = 0.0, sres = 0.0;
      for (long i = 0; i < m; ++i) {
        const double logits = tmp[i] + x[d];
        const double u = 1.0 + exp(-logits);
        const double yi = p->fvec[i];
        val += (yi - 1.0) * logits - log(u);                  /* -mean((y - 1) .* logits - log.(u)) */
        tmp[i] = 1.0 / u - yi;                                /* probs - y */
        sres += tmp[i];
      }
      amul(p->F, m, d, tmp, grad);
      for (long j = 0; j < d; ++j) grad[j] /= (double)m;
      grad[d] = sres / (double)m;
      return -(val / (double)m);
    }
    case F_QUADRATIC: {                                       /* temp = Q x ; 0.5 dot(x, temp) + dot(x, q) ; temp + q */
      mul(p->F, n, n, x, tmp);
      const double v = 0.5 * dot(x, tmp, n) + dot(x, p->fvec, n);
      for (long i = 0; i < n; ++i) grad[i] = tmp[i] + p->fvec[i];
      return v;
    }
    default:                                                  /* Zero: 0, zero(x) */
      for (long i = 0; i < n; ++i) grad[i] = 0.0;
      return 0.0;
  }
}

/* ---- prox operators ----------------------------------------------------------------------------------------------- */
/* y = prox_{gamma f}(x); Translate: prox of f at x + b, minus b */
static void prox(const ref_prox* f, const double* x, double gamma, double* y, long n) {
  double scale = 1.0;
  if (f->kind == P_NORM_L2) {
    double s = 0.0;
    for (long i = 0; i < n; ++i) { const double z = f->shift ? x[i] + f->shift[i] : x[i]; s += z * z; }
    scale = 1.0 - f->lambda * gamma / sqrt(s);
    if (!(scale > 0.0)) scale = 0.0;                          /* max(0, .) */
  }
  for (long i = 0; i < n; ++i) {
    const double z = f->shift ? x[i] + f->shift[i] : x[i];
    double v;
    switch (f->kind) {
      case P_ZERO: v = z; break;
      case P_IND_ZERO: v = 0.0; break;
      case P_NORM_L1: { const double gl = gamma * f->lambda; v = z + (z <= -gl ? gl : (z >= gl ? -gl : -z)); } break;
      case P_NORM_L2: v = scale * z; break;
      default: v = z < f->lo ? f->lo : (z > f->hi ? f->hi : z); break;
    }
    y[i] = f->shift ? v - f->shift[i] : v;
  }
}
/* y = prox_{sigma f*}(w) by Moreau, ProximalCore's order: y = prox_{f / sigma}(w / sigma); out = w - sigma y.
 * convex_conjugate(Zero) = IndZero and convex_conjugate(IndZero) = Zero are direct (no shift). */
static void prox_conj(const ref_prox* f, const double* w, double sigma, double* y, double* tmp, long n) {
  if (!f->shift && f->kind == P_ZERO) { for (long i = 0; i < n; ++i) y[i] = 0.0; return; }
  if (!f->shift && f->kind == P_IND_ZERO) { for (long i = 0; i < n; ++i) y[i] = w[i]; return; }
  for (long i = 0; i < n; ++i) tmp[i] = w[i] / sigma;
  prox(f, tmp, 1.0 / sigma, y, n);
  for (long i = 0; i < n; ++i) y[i] = w[i] - sigma * y[i];
}
static double prox_value(const ref_prox* f, const double* x, long n) {
  double s = 0.0;
  for (long i = 0; i < n; ++i) {
    const double z = f->shift ? x[i] + f->shift[i] : x[i];
    switch (f->kind) {
      case P_ZERO: break;
      case P_IND_ZERO: if (z != 0.0) return INFINITY; break;
      case P_NORM_L1: s += fabs(z); break;
      case P_NORM_L2: s += z * z; break;
      default: if (z < f->lo || z > f->hi) return INFINITY; break;
    }
  }
  if (f->kind == P_NORM_L1) return f->lambda * s;
  if (f->kind == P_NORM_L2) return f->lambda * sqrt(s);
  return 0.0;
}

/* ---- stepsize rules (:208-273); state = (s0, s1) ------------------------------------------------------------------- */
static void rule_init(const ref_problem* p, double* gamma, double* sigma, double* s0, double* s1) {
  *gamma = p->gamma;
  *sigma = p->gamma * (p->t * p->t);
  if (p->rule == RULE_MM) { *s0 = p->gamma; *s1 = INFINITY; }
  else { *s0 = p->gamma; *s1 = p->gamma; }
}
static void rule_step(const ref_problem* p, const double* x1, const double* g1, const double* x0, const double* g0, long n,
                      double* dgr, double* dx, double* gamma, double* sigma, double* s0, double* s1) {
  if (p->rule == RULE_FIXED) { *gamma = p->gamma; *sigma = p->gamma * (p->t * p->t); return; }
  for (long i = 0; i < n; ++i) { dgr[i] = g1[i] - g0[i]; dx[i] = x1[i] - x0[i]; }
  if (p->rule == RULE_MM) {                                                       /* :226-230 */
    const double gamma_prev = *s0, rho = *s1;
    const double L = norm2(dgr, n) / norm2(dx, n);
    const double g = jl_min(sqrt(1.0 + rho) * gamma_prev, 1.0 / (2.0 * L));
    *gamma = g; *sigma = g * (p->t * p->t); *s0 = g; *s1 = g / gamma_prev;
    return;
  }
  const double gamma1 = *s0, gamma0 = *s1;                                        /* :258-273 */
  const double xi = (p->t * p->t) * (gamma1 * gamma1) * (p->norm_A * p->norm_A);
  const double dgx = dot(dgr, dx, n);
  const double C = nan_to_zero(sq(norm2(dgr, n)) / dgx);
  const double L = nan_to_zero(dgx / sq(norm2(dx, n)));
  const double D = gamma1 * L * (gamma1 * C - 1.0);
  const double d1 = 1.0 + p->delta;
  const double m4 = 1.0 - 4.0 * xi * (d1 * d1);
  const double g = jl_min(jl_min(gamma1 * sqrt(1.0 + gamma1 / gamma0), 1.0 / (2.0 * p->Theta * p->t * p->norm_A)),
                          gamma1 * sqrt(m4) / sqrt(2.0 * d1 * (D + sqrt(D * D + xi * m4))));
  *gamma = g; *sigma = g * (p->t * p->t); *s0 = g; *s1 = gamma1;
}

/* ---- adaptive_primal_dual (:312-364) -------------------------------------------------------------------------------
 * hist arrays (may be NULL) receive gamma, sigma, norm_res, objective of iterations 1..min(it, nhist).
 * Returns the reference's `it` (maxit when the tolerance was not reached).  y0 / y_out have length am (n when A is NULL). */
long ref_adaptive_primal_dual(const ref_problem* p, const double* x0, const double* y0, double* x_out, double* y_out,
                              double* gamma_hist, double* sigma_hist, double* res_hist, double* obj_hist, long nhist) {
  const long n = p->n, md = p->A ? p->am : n;
  const long big = (p->fm > n ? p->fm : n) > md ? (p->fm > n ? p->fm : n) : md;
  double* buf = (double*)calloc((size_t)(9 * n + 6 * md + 2 * big), sizeof(double));
  if (!buf) return -1;
  double *x = buf, *x_prev = x + n, *grad = x_prev + n, *grad_prev = grad + n, *v = grad_prev + n, *At_y = v + n,
         *pres = At_y + n, *dgr = pres + n, *dx = dgr + n;
  double *y = dx + n, *A_x = y + md, *A_x_prev = A_x + md, *w = A_x_prev + md, *dres = w + md, *ynew = dres + md;
  double *tmp = ynew + md, *tmp2 = tmp + big;
  memcpy(x, x0, (size_t)n * sizeof(double));
  if (y0) memcpy(y, y0, (size_t)md * sizeof(double));

  double gamma, sigma, s0, s1;
  rule_init(p, &gamma, &sigma, &s0, &s1);                                         /* :324 */
  /* h_conj = convex_conjugate(h) is applied through prox_conj                       :325 */
  if (p->A) mul(p->A, md, n, x, A_x); else for (long i = 0; i < md; ++i) A_x[i] = 0.0 * x[i];          /* :327 */
  eval_f(p, x, grad, tmp);                                                        /* :328 */
  if (p->A) amul(p->A, md, n, y, At_y); else for (long i = 0; i < n; ++i) At_y[i] = 0.0 * y[i];        /* :329 */
  for (long i = 0; i < n; ++i) v[i] = x[i] - gamma * (grad[i] + At_y[i]);         /* :330 */
  memcpy(x_prev, x, (size_t)n * sizeof(double));                                  /* :331 */
  memcpy(A_x_prev, A_x, (size_t)md * sizeof(double));
  memcpy(grad_prev, grad, (size_t)n * sizeof(double));
  prox(&p->g, v, gamma, x, n);                                                    /* :332 */

  long it_ret = p->maxit;
  for (long it = 1; it <= p->maxit; ++it) {
    if (p->A) mul(p->A, md, n, x, A_x); else for (long i = 0; i < md; ++i) A_x[i] = 0.0 * x[i];        /* :335 */
    const double f_x = eval_f(p, x, grad, tmp);                                   /* :336 */
    for (long i = 0; i < n; ++i) pres[i] = (v[i] - x[i]) / gamma + grad[i] + At_y[i];                  /* :338 */
    const double gamma_prev = gamma;                                              /* :340 */
    rule_step(p, x, grad, x_prev, grad_prev, n, dgr, dx, &gamma, &sigma, &s0, &s1);                    /* :341 */
    const double rho = gamma / gamma_prev;                                        /* :342 */
    for (long i = 0; i < md; ++i) w[i] = y[i] + sigma * ((1.0 + rho) * A_x[i] - rho * A_x_prev[i]);    /* :344 */
    prox_conj(&p->h, w, sigma, ynew, tmp2, md);                                   /* :345 */
    memcpy(y, ynew, (size_t)md * sizeof(double));
    for (long i = 0; i < md; ++i) dres[i] = (w[i] - y[i]) / sigma - A_x[i];       /* :347 */
    const double norm_res = sqrt(sq(norm2(pres, n)) + sq(norm2(dres, md)));       /* :348 */
    if (it <= nhist) {                                                            /* :350-352 */
      if (gamma_hist) gamma_hist[it - 1] = gamma;
      if (sigma_hist) sigma_hist[it - 1] = sigma;
      if (res_hist) res_hist[it - 1] = norm_res;
      if (obj_hist) obj_hist[it - 1] = f_x + prox_value(&p->g, x, n) + (p->A ? prox_value(&p->h, A_x, md) : 0.0);
    }
    if (norm_res <= p->tol) { it_ret = it; break; }                               /* :354-356 */
    if (p->A) amul(p->A, md, n, y, At_y); else for (long i = 0; i < n; ++i) At_y[i] = 0.0 * y[i];      /* :358 */
    for (long i = 0; i < n; ++i) v[i] = x[i] - gamma * (grad[i] + At_y[i]);       /* :359 */
    memcpy(x_prev, x, (size_t)n * sizeof(double));                                /* :360 */
    memcpy(A_x_prev, A_x, (size_t)md * sizeof(double));
    memcpy(grad_prev, grad, (size_t)n * sizeof(double));
    prox(&p->g, v, gamma, x, n);                                                  /* :361 */
  }
  memcpy(x_out, x, (size_t)n * sizeof(double));
  if (y_out) memcpy(y_out, y, (size_t)md * sizeof(double));
  free(buf);
  return it_ret;
}
