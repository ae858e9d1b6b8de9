/* adaprox_ref.c -- TEST INFRASTRUCTURE: a second, independent CPU restatement of the reference's generic loop in plain C.
 *
 * Written directly from the Julia source (not from oracle/adaprox_oracle.py) so that the numpy oracle -- the checker of every
 * GPU parity test -- is itself cross-checked by an implementation that shares no code, no BLAS and no summation order with it
 * (tests/test_oracle_c_restatement.py).  Only tests/ may load this library; the product never does.
 *
 * Follows, line by line:
 *   adaptive_primal_dual            src/AdaProx.jl:312-364   (AdaPGM :418-421 = the same loop with h = Zero(), A = 0, y = zero(x))
 *   adaptive_linesearch_primal_dual src/AdaProx.jl:463-550   (AdaPDM+)
 *   backtrack_stepsize, backtracking_proxgrad, backtracking_nesterov :34-84 ; fixed_nesterov :91-142 ; agraal :150-192
 *   backtrack_stepsize_MP, malitsky_pock                      src/AdaProx.jl:555-629
 *   FixedStepsize / MalitskyMishchenkoRule / OurRule / OurRulePlus   src/AdaProx.jl:208-308
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
enum { RULE_FIXED = 0, RULE_MM = 1, RULE_OUR = 2, RULE_OUR_PLUS = 3 };

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
  double xi, nu, r;             /* OurRulePlus (:277-308) */
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
  *sigma = (p->rule == RULE_OUR_PLUS) ? p->gamma : p->gamma * (p->t * p->t);     /* :294-297: (gamma, gamma) */
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
  const double gamma1 = *s0, gamma0 = *s1;
  if (p->rule == RULE_OUR_PLUS) {                                                 /* :299-308 */
    const double dgx2 = dot(dgr, dx, n);
    const double C2 = nan_to_zero(sq(norm2(dgr, n)) / dgx2);
    const double L2 = nan_to_zero(dgx2 / sq(norm2(dx, n)));
    const double D2 = nan_to_zero(1.0 - 2.0 * p->r + gamma1 * L2 * (gamma1 * C2 + 2.0 * (p->r - 1.0)));
    const double Dpos = D2 > 0.0 ? D2 : 0.0;                                      /* max(D, 0); D is not NaN here */
    const double g2 = gamma1 * jl_min(sqrt(1.0 / (p->r * (p->nu + p->xi)) + gamma1 / gamma0),
                                      sqrt((p->nu * (1.0 + p->xi) - 1.0) / (p->nu * (p->nu + p->xi))) / sqrt(Dpos));
    *gamma = g2; *sigma = g2; *s0 = g2; *s1 = gamma1;
    return;
  }
  const double xi = (p->t * p->t) * (gamma1 * gamma1) * (p->norm_A * p->norm_A);  /* :258-273 */
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

/* ---- adaptive_linesearch_primal_dual, AdaPDM+ (:463-550) -------------------------------------------------------------
 * p->gamma must already be resolved (gamma === nothing -> 1 / (2 Theta t eta), :484-486); rule fields other than
 * t, delta, Theta are unused.  eta, r, R as in the reference.  trials_out (may be NULL): total linesearch trials. */
long ref_adaptive_linesearch_primal_dual(const ref_problem* p, double eta, double r, double R, const double* x0, const double* y0,
                                         double* x_out, double* y_out, double* gamma_hist, double* sigma_hist, double* res_hist,
                                         double* obj_hist, long nhist, long* trials_out) {
  const long n = p->n, md = p->am;
  if (!p->A) return -2;
  const long big = (p->fm > n ? p->fm : n) > md ? (p->fm > n ? p->fm : n) : md;
  double* buf = (double*)calloc((size_t)(10 * n + 6 * md + 2 * big), sizeof(double));
  if (!buf) return -1;
  double *x = buf, *x_prev = x + n, *grad = x_prev + n, *grad_prev = grad + n, *v = grad_prev + n, *At_y = v + n,
         *pres = At_y + n, *dgr = pres + n, *dx = dgr + n, *At_y_next = dx + n;
  double *y = At_y_next + n, *A_x = y + md, *A_x_prev = A_x + md, *w = A_x_prev + md, *dres = w + md, *y_next = dres + md;
  double *tmp = y_next + md, *tmp2 = tmp + big;
  memcpy(x, x0, (size_t)n * sizeof(double));
  memcpy(y, y0, (size_t)md * sizeof(double));
  const double t = p->t, Theta = p->Theta, delta1 = 1.0 + p->delta;               /* :490 */
  double gamma = p->gamma, gamma_prev = gamma, sigma = t * t * gamma;             /* :491 */
  long trials = 0;
  mul(p->A, md, n, x, A_x);                                                       /* :494 */
  eval_f(p, x, grad, tmp);
  amul(p->A, md, n, y, At_y);
  for (long i = 0; i < n; ++i) v[i] = x[i] - gamma * (grad[i] + At_y[i]);         /* :497 */
  memcpy(x_prev, x, (size_t)n * sizeof(double));
  memcpy(A_x_prev, A_x, (size_t)md * sizeof(double));
  memcpy(grad_prev, grad, (size_t)n * sizeof(double));
  prox(&p->g, v, gamma, x, n);                                                    /* :499 */
  long it_ret = p->maxit;
  for (long it = 1; it <= p->maxit; ++it) {
    mul(p->A, md, n, x, A_x);                                                     /* :502 */
    const double f_x = eval_f(p, x, grad, tmp);
    for (long i = 0; i < n; ++i) pres[i] = (v[i] - x[i]) / gamma + grad[i] + At_y[i];                  /* :505 */
    for (long i = 0; i < n; ++i) { dgr[i] = grad[i] - grad_prev[i]; dx[i] = x[i] - x_prev[i]; }
    const double dgx = dot(dgr, dx, n);
    const double C = nan_to_zero(sq(norm2(dgr, n)) / dgx);                        /* :507 */
    const double L = nan_to_zero(dgx / sq(norm2(dx, n)));                         /* :508 */
    const double Delta = gamma * L * (gamma * C - 1.0);                           /* :509 */
    const double xi_bar = (t * t) * (gamma * gamma) * (eta * eta) * (delta1 * delta1);                 /* :510 */
    const double m4xim1 = 1.0 - 4.0 * xi_bar;                                     /* :511 */
    eta = R * eta;                                                                /* :513 */
    for (;;) {                                                                    /* :516-533 */
      ++trials;
      const double gamma_next = jl_min(jl_min(gamma * sqrt(1.0 + gamma / gamma_prev), 1.0 / (2.0 * Theta * t * eta)),
                                       gamma * sqrt(m4xim1 / (2.0 * delta1 * (Delta + sqrt(Delta * Delta + m4xim1 * sq(t * eta * gamma))))));
      const double rho = gamma_next / gamma;
      sigma = (t * t) * gamma_next;
      for (long i = 0; i < md; ++i) w[i] = y[i] + sigma * ((1.0 + rho) * A_x[i] - rho * A_x_prev[i]);
      prox_conj(&p->h, w, sigma, y_next, tmp2, md);
      amul(p->A, md, n, y_next, At_y_next);
      double dn = 0.0, dd = 0.0;
      for (long j = 0; j < n; ++j) dn += sq(At_y_next[j] - At_y[j]);
      for (long i = 0; i < md; ++i) dd += sq(y_next[i] - y[i]);
      if (eta >= sqrt(dn) / sqrt(dd)) {                                           /* :527 */
        gamma_prev = gamma; gamma = gamma_next;
        memcpy(y, y_next, (size_t)md * sizeof(double));
        memcpy(At_y, At_y_next, (size_t)n * sizeof(double));
        break;
      }
      eta *= r;                                                                   /* :532 */
      if (trials > 100000000L) break;
    }
    for (long i = 0; i < md; ++i) dres[i] = (w[i] - y[i]) / sigma - A_x[i];       /* :535 */
    const double norm_res = sqrt(sq(norm2(pres, n)) + sq(norm2(dres, md)));
    if (it <= nhist) {
      if (gamma_hist) gamma_hist[it - 1] = gamma;
      if (sigma_hist) sigma_hist[it - 1] = sigma;
      if (res_hist) res_hist[it - 1] = norm_res;
      if (obj_hist) obj_hist[it - 1] = f_x + prox_value(&p->g, x, n) + prox_value(&p->h, A_x, md);
    }
    if (norm_res <= p->tol) { it_ret = it; break; }                               /* :541-543 */
    for (long i = 0; i < n; ++i) v[i] = x[i] - gamma * (grad[i] + At_y[i]);       /* :545 */
    memcpy(x_prev, x, (size_t)n * sizeof(double));
    memcpy(A_x_prev, A_x, (size_t)md * sizeof(double));
    memcpy(grad_prev, grad, (size_t)n * sizeof(double));
    prox(&p->g, v, gamma, x, n);                                                  /* :547 */
  }
  memcpy(x_out, x, (size_t)n * sizeof(double));
  memcpy(y_out, y, (size_t)md * sizeof(double));
  if (trials_out) *trials_out = trials;
  free(buf);
  return it_ret;
}

/* ---- the proximal-gradient baselines --------------------------------------------------------------------------------
 * which: 0 backtracking_proxgrad (:50-64), 1 backtracking_nesterov (:66-84), 2 fixed_nesterov (:91-142, mu = 0 or > 0),
 *        3 agraal (:150-192; x_second = the other start point x0, gamma0 <= 0 means `nothing`).
 * p->gamma = gamma0 (0, 1, 3) / gamma (2).  evals_out[0..1] (may be NULL): f evaluations, gradient evaluations as
 * Counting would report them (the logged f(x) of fixed_nesterov / agraal is not counted). */
static double upper_bound(const double* x, double f_x, const double* grad_x, const double* z, double gamma, long n) {   /* :26 */
  double gd = 0.0, dd = 0.0;
  for (long i = 0; i < n; ++i) { const double d = z[i] - x[i]; gd += grad_x[i] * d; dd += d * d; }
  return f_x + gd + 1.0 / (2.0 * gamma) * sq(sqrt(dd));
}
long ref_proxgrad_family(const ref_problem* p, int which, double xi, double shrink, double muf, double mug, double theta0,
                         double gamma_max, double phi, const double* x0, const double* x_second, double* x_out,
                         double* gamma_hist, double* res_hist, double* obj_hist, long nhist, long* evals_out) {
  const long n = p->n;
  const long big = p->fm > n ? p->fm : n;
  double* buf = (double*)calloc((size_t)(8 * n + big), sizeof(double));
  if (!buf) return -1;
  double *x = buf, *z = x + n, *z_prev = z + n, *grad = z_prev + n, *grad2 = grad + n, *u = grad2 + n, *x_bar = u + n, *x_prev = x_bar + n;
  double* tmp = x_prev + n;
  long n_eval = 0, n_grad = 0, it_ret = p->maxit;
  double gamma = p->gamma;
  memcpy(x, x0, (size_t)n * sizeof(double));
  double* result = x;
#define HIST(it, g_, r_, o_) do { if ((it) <= nhist) { if (gamma_hist) gamma_hist[(it) - 1] = (g_); if (res_hist) res_hist[(it) - 1] = (r_); \
                                                        if (obj_hist) obj_hist[(it) - 1] = (o_); } } while (0)
  if (which == 0 || which == 1) {
    memcpy(z, x, (size_t)n * sizeof(double));
    double theta = 1.0;
    double f_x = eval_f(p, x, grad, tmp); n_eval++; n_grad++;                     /* :52 / :69 */
    result = z;
    for (long it = 1; it <= p->maxit; ++it) {
      if (which == 1) memcpy(z_prev, z, (size_t)n * sizeof(double));              /* :71 */
      gamma = (which == 0) ? xi * gamma : gamma;                                  /* :54 / :72 */
      double f_z;
      for (;;) {                                                                  /* backtrack_stepsize :34-48 */
        for (long i = 0; i < n; ++i) u[i] = x[i] - gamma * grad[i];
        prox(&p->g, u, gamma, z, n);
        const double ub_z = upper_bound(x, f_x, grad, z, gamma, n);
        f_z = eval_f(p, z, grad2, tmp); n_eval++;                                 /* value now, pullback (grad2) on demand */
        if (!(f_z > ub_z)) break;
        gamma *= shrink;
        if (gamma < 1e-300) break;
      }
      double dd = 0.0;
      for (long i = 0; i < n; ++i) dd += sq(z[i] - x[i]);
      const double norm_res = sqrt(dd) / gamma;                                   /* :55 / :73 */
      HIST(it, gamma, norm_res, f_z + prox_value(&p->g, z, n));
      if (norm_res <= p->tol) { it_ret = it; break; }
      if (which == 0) {                                                           /* :60-61 */
        memcpy(x, z, (size_t)n * sizeof(double)); f_x = f_z;
        memcpy(grad, grad2, (size_t)n * sizeof(double)); n_grad++;
      } else {                                                                    /* :78-81 */
        const double theta_prev = theta;
        theta = (1.0 + sqrt(1.0 + 4.0 * theta_prev * theta_prev)) / 2.0;
        for (long i = 0; i < n; ++i) x[i] = z[i] + (theta_prev - 1.0) / theta * (z[i] - z_prev[i]);
        f_x = eval_f(p, x, grad, tmp); n_eval++; n_grad++;
      }
    }
  } else if (which == 2) {
    const double mu = muf + mug;                                                  /* :108-117 */
    const double q = gamma * mu / (1.0 + gamma * mug);
    double theta = theta0 >= 0.0 ? theta0 : (q > 0.0 ? 1.0 / sqrt(q) : 0.0);
    memcpy(x_prev, x, (size_t)n * sizeof(double));                                /* :119 */
    for (long it = 1; it <= p->maxit; ++it) {
      const double theta_prev = theta;
      double beta;
      if (mu == 0.0) {                                                            /* :122-128 */
        theta = (1.0 + sqrt(1.0 + 4.0 * theta_prev * theta_prev)) / 2.0;
        beta = (theta_prev - 1.0) / theta;
      } else {
        const double a = 1.0 - q * theta_prev * theta_prev;
        theta = (a + sqrt(a * a + 4.0 * theta_prev * theta_prev)) / 2.0;
        beta = (theta_prev - 1.0) * (1.0 + gamma * mug - theta * gamma * mu) / theta / (1.0 - gamma * muf);
      }
      for (long i = 0; i < n; ++i) z[i] = x[i] + beta * (x[i] - x_prev[i]);       /* :129 */
      eval_f(p, z, grad, tmp); n_eval++; n_grad++;                                /* :130 */
      memcpy(x_prev, x, (size_t)n * sizeof(double));                              /* :131 */
      for (long i = 0; i < n; ++i) u[i] = z[i] - gamma * grad[i];
      prox(&p->g, u, gamma, x, n);                                                /* :132 */
      double dd = 0.0;
      for (long i = 0; i < n; ++i) dd += sq(x[i] - z[i]);
      const double norm_res = sqrt(dd) / gamma;                                   /* :133 */
      if (it <= nhist && obj_hist) { const double fx = eval_f(p, x, grad2, tmp); HIST(it, gamma, norm_res, fx + prox_value(&p->g, x, n)); }
      else HIST(it, gamma, norm_res, NAN);
      if (norm_res <= p->tol) { it_ret = it; break; }
    }
  } else {
    memcpy(x_prev, x_second, (size_t)n * sizeof(double));                         /* :165 */
    memcpy(x_bar, x, (size_t)n * sizeof(double));
    eval_f(p, x, grad, tmp);                                                      /* :166 */
    eval_f(p, x_prev, grad2, tmp);                                                /* :167 */
    n_eval = 2; n_grad = 2;
    const double rho = 1.0 / phi + 1.0 / (phi * phi);                             /* :172 */
    double theta = 1.0;
    for (long it = 1; it <= p->maxit; ++it) {
      double dxx = 0.0, dgg = 0.0;
      for (long i = 0; i < n; ++i) { dxx += sq(x[i] - x_prev[i]); dgg += sq(grad[i] - grad2[i]); }
      if (it == 1 && !(p->gamma > 0.0)) gamma = sqrt(dxx) / sqrt(dgg);            /* :168-170 */
      const double C = sq(sqrt(dxx)) / sq(sqrt(dgg));                             /* :175 */
      const double gamma_prev = gamma;
      gamma = jl_min(jl_min(rho * gamma_prev, phi * theta * C / (4.0 * gamma_prev)), gamma_max);       /* :177 */
      theta = phi * gamma / gamma_prev;                                           /* :178 */
      for (long i = 0; i < n; ++i) x_bar[i] = ((phi - 1.0) * x[i] + x_bar[i]) / phi;                   /* :179 */
      memcpy(x_prev, x, (size_t)n * sizeof(double));                              /* :180 */
      memcpy(grad2, grad, (size_t)n * sizeof(double));
      for (long i = 0; i < n; ++i) u[i] = x_bar[i] - gamma * grad2[i];
      prox(&p->g, u, gamma, x, n);                                                /* :181 */
      double dd = 0.0;
      for (long i = 0; i < n; ++i) dd += sq(x[i] - x_prev[i]);
      const double norm_res = sqrt(dd) / gamma;                                   /* :182 */
      const double fx = eval_f(p, x, grad, tmp);                                  /* value for the record (uncounted) + :189's gradient */
      HIST(it, gamma, norm_res, fx + prox_value(&p->g, x, n));
      if (norm_res <= p->tol) { it_ret = it; break; }
      n_eval++; n_grad++;                                                         /* :189 */
    }
  }
#undef HIST
  memcpy(x_out, result, (size_t)n * sizeof(double));
  if (evals_out) { evals_out[0] = n_eval; evals_out[1] = n_grad; }
  free(buf);
  return it_ret;
}

/* ---- malitsky_pock (:555-629) ------------------------------------------------------------------------------------------ */
long ref_malitsky_pock(const ref_problem* p, double sigma, const double* x0, const double* y0, double* x_out, double* y_out,
                       double* gamma_hist, double* sigma_hist, double* res_hist, double* obj_hist, long nhist) {
  const long n = p->n, md = p->am;
  if (!p->A) return -2;
  const long big = (p->fm > n ? p->fm : n) > md ? (p->fm > n ? p->fm : n) : md;
  double* buf = (double*)calloc((size_t)(8 * n + 5 * md + 2 * big), sizeof(double));
  if (!buf) return -1;
  double *x = buf, *x_prev = x + n, *grad = x_prev + n, *grad_prev = grad + n, *v = grad_prev + n, *At_y = v + n, *At_y_prev = At_y + n,
         *pres = At_y_prev + n;
  double *y = pres + n, *A_x = y + md, *A_x_prev = A_x + md, *w = A_x_prev + md, *dres = w + md;
  double *tmp = dres + md, *tmp2 = tmp + big;
  memcpy(x, x0, (size_t)n * sizeof(double));
  memcpy(y, y0, (size_t)md * sizeof(double));
  const double t = p->t, theta1 = 1.0;                                            /* :595: theta = one(sigma), never updated */
  mul(p->A, md, n, x, A_x);                                                       /* :597 */
  amul(p->A, md, n, y, At_y);                                                     /* :598 */
  long it_ret = p->maxit;
  for (long it = 1; it <= p->maxit; ++it) {
    memcpy(At_y_prev, At_y, (size_t)n * sizeof(double));                          /* :600 */
    for (long i = 0; i < md; ++i) w[i] = y[i] + sigma * A_x[i];                   /* :601 */
    prox_conj(&p->h, w, sigma, y, tmp2, md);                                      /* :602 (w and y are distinct buffers) */
    amul(p->A, md, n, y, At_y);                                                   /* :603 */
    const double sigma_prev = sigma;                                              /* :605-606 */
    sigma = sigma * sqrt(1.0 + theta1);
    const double f_x_prev = eval_f(p, x, grad_prev, tmp);                         /* :608 */
    memcpy(x_prev, x, (size_t)n * sizeof(double));                                /* :609 */
    memcpy(A_x_prev, A_x, (size_t)md * sizeof(double));
    double gamma, f_x;
    for (;;) {                                                                    /* backtrack_stepsize_MP :555-579 */
      const double th = sigma / sigma_prev;
      gamma = t * t * sigma;
      for (long i = 0; i < n; ++i) v[i] = x_prev[i] - gamma * (((1.0 + th) * At_y[i] - th * At_y_prev[i]) + grad_prev[i]);
      prox(&p->g, v, gamma, x, n);
      mul(p->A, md, n, x, A_x);
      f_x = eval_f(p, x, grad, tmp);
      double da = 0.0, dxx = 0.0, gd = 0.0;
      for (long i = 0; i < md; ++i) da += sq(A_x[i] - A_x_prev[i]);
      for (long i = 0; i < n; ++i) { const double d = x[i] - x_prev[i]; dxx += d * d; gd += grad_prev[i] * d; }
      const double lhs = gamma * sigma * sq(sqrt(da)) + 2.0 * gamma * (f_x - f_x_prev - gd);
      if (!(lhs > 0.95 * sq(sqrt(dxx)))) break;
      sigma /= 2.0;
      if (sigma < 1e-300) break;
    }
    for (long i = 0; i < n; ++i) pres[i] = (v[i] - x[i]) / gamma + grad[i] + At_y[i];                  /* :616 */
    for (long i = 0; i < md; ++i) dres[i] = (w[i] - y[i]) / sigma_prev - A_x[i];                       /* :617 */
    const double norm_res = sqrt(sq(norm2(pres, n)) + sq(norm2(dres, md)));
    if (it <= nhist) {
      if (gamma_hist) gamma_hist[it - 1] = gamma;
      if (sigma_hist) sigma_hist[it - 1] = sigma;
      if (res_hist) res_hist[it - 1] = norm_res;
      if (obj_hist) obj_hist[it - 1] = f_x + prox_value(&p->g, x, n) + prox_value(&p->h, A_x, md);
    }
    if (norm_res <= p->tol) { it_ret = it; break; }
  }
  memcpy(x_out, x, (size_t)n * sizeof(double));
  memcpy(y_out, y, (size_t)md * sizeof(double));
  free(buf);
  return it_ret;
}

/* ---- ProximalCore.prox(f, x, gamma) -> (y, f(y)) and the same through convex_conjugate(f), one call each (tests) ------- */
double ref_prox_eval(const ref_prox* f, int conjugate, const double* x, double gamma, double* y, long n) {
  if (conjugate) {
    double* tmp = (double*)malloc((size_t)(n > 0 ? n : 1) * sizeof(double));
    if (!tmp) return NAN;
    prox_conj(f, x, gamma, y, tmp, n);
    free(tmp);
    return NAN;                              /* the solvers never use the value of a conjugate (`y, _ = prox(...)`) */
  }
  prox(f, x, gamma, y, n);
  return prox_value(f, y, n);
}
