// phases.cuh -- prox operators, stepsize rules and the smooth-oracle phases.
//
// Everything here is a grid-wide "phase": a device function that every thread
// of every CTA of the persistent grid calls, followed (by the caller) by a
// grid sync.  Reductions leave per-CTA partials in W.red[slot][cta]; the next
// phase sums them in a fixed order (grid_totals), so all CTAs agree bit for bit
// on every scalar without a single-thread bottleneck or a host round trip.
#pragma once
#include "gemv.cuh"
#include "p2p_args.cuh"
#include "../../include/adaprox.h"

namespace adaprox {

// reduction slots
enum {
  SLOT_F0 = 0, SLOT_F1 = 1,            // smooth-term value sums
  SLOT_XX0 = 2,                        // |x|^2 (cubic)
  SLOT_PR = 3, SLOT_GG = 4, SLOT_GX = 5, SLOT_DXX = 6,   // |primal_res|^2, |dgrad|^2, <dgrad,dx>, |dx|^2
  SLOT_GVAL = 7,                       // g(x) sum
  SLOT_DR = 8, SLOT_HVAL = 9, SLOT_L2 = 10,              // |dual_res|^2, h(Ax) sum, |z|^2 for the NormL2 prox
  SLOT_DY = 11, SLOT_DATY = 12,        // AdaPDM+: |y+ - y|^2, |A'y+ - A'y|^2
  SLOT_AUX0 = 13, SLOT_AUX1 = 14, SLOT_AUX2 = 15
};

struct DProblem {
  int f_kind, f_ipar;
  double f_c;
  DMat F;                  // matrix of the smooth term
  const double* fvec;      // b / y / q (local rows for LS / logistic; full length for Quadratic)
  int64_t f_row0;          // Quadratic with a row-sharded Q: F holds rows [f_row0, f_row0 + F.m) of the n x n matrix
  double f_N;              // logistic: global number of samples
  DProx g, h;
  DMat A;                  // linear map of the primal-dual solvers (MAT_NONE: `A = 0`)
  int64_t n, md;           // primal / dual dimension (md: LOCAL rows of A when A is a row shard)
  P2PArgs p2p;             // row-sharded primal-dual solves: in-kernel all-reduce over peer memory (p2p.n <= 1: single GPU)
  int A_sharded, F_sharded;   // which of the two matrices is a row block (p2p.n > 1 only)
};

struct DOpts {
  int solver, rule;
  double gamma, t, norm_A, delta, Theta, xi, nu, r, R, eta, shrink, sigma, muf, mug, theta, gamma_max, phi, tol;
  int64_t maxit, max_records;
  int want_objective;
};

struct DResult {
  int64_t iters;
  unsigned flags;
  int xbuf;                // unused
  int64_t f_evals, grad_f_evals, prox_g_evals, prox_h_evals, A_evals, At_evals, n_records;
  double final_gamma, final_sigma, final_norm_res;
};

struct DWork {
  double* xb[3];           // iterate ring
  double* gb[2];           // gradient ring
  double* v; double* Aty[2]; double* yb[2]; double* w; double* Axb[2]; double* r;
  double* aux[3];          // extra n-vectors (proximal-gradient family)
  double* fu;              // Gram-form Quadratic: u = Z'x (F.n entries), see phases_pre.cuh
  double* red;             // [kMaxRed][G]
  double* xout; double* yout;
  adaprox_record* rec;
  DResult* res;
  unsigned long long* tstamp;   // optional [iters][8] globaltimer stamps at phase boundaries (ADAPROX_PHASE_TIMING=1)
  int tstamp_iters;
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// CTA 0 / thread 0 stamps the time at which it passed phase boundary `k` of iteration `it`
__device__ __forceinline__ void phase_stamp(const DWork& W, long long it, int k) {
  if (W.tstamp != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && it >= 1 && it <= W.tstamp_iters)
    W.tstamp[(it - 1) * 8 + k] = globaltimer_ns();
}

// ---------------------------------------------------------------------------
// prox operators (ProximalCore / ProximalOperators bodies)
// ---------------------------------------------------------------------------
// prox_{gamma f}(x)_i for the separable kinds; NormL2 uses the precomputed scale.
__device__ __forceinline__ double prox_elem(const DProx& p, double x, double gamma, int64_t i, double l2scale) {
  const double s = p.shift ? p.shift[i] : 0.0;
  const double z = p.shift ? x + s : x;                   // Translate: z = x + b
  double y;
  switch (p.kind) {
    case ADAPROX_P_ZERO: y = z; break;
    case ADAPROX_P_IND_ZERO: y = 0.0; break;
    case ADAPROX_P_NORM_L1: {
      const double gl = gamma * p.lambda;
      y = z + (z <= -gl ? gl : (z >= gl ? -gl : -z));
    } break;
    case ADAPROX_P_NORM_L2: y = l2scale * z; break;
    default: {  // IND_BOX
      const double lo = p.lo_vec ? p.lo_vec[i] : p.lo, hi = p.hi_vec ? p.hi_vec[i] : p.hi;
      y = z < lo ? lo : (z > hi ? hi : z);
    } break;
  }
  return p.shift ? y - s : y;                              // y .-= b
}

// prox_{sigma f*}(w)_i through Moreau, in ProximalCore's order:
//   u = w / sigma ; y = prox_{f / sigma}(u) ; out = w - sigma * y
// convex_conjugate(Zero) = IndZero and convex_conjugate(IndZero) = Zero are direct.
__device__ __forceinline__ double prox_conj_elem(const DProx& p, double w, double sigma, int64_t i, double l2scale) {
  if (!p.shift && p.kind == ADAPROX_P_ZERO) return 0.0;
  if (!p.shift && p.kind == ADAPROX_P_IND_ZERO) return w;
  const double u = w / sigma;
  const double y = prox_elem(p, u, 1.0 / sigma, i, l2scale);
  return w - sigma * y;
}

// the argument whose 2-norm the NormL2 prox needs: z = x + shift
__device__ __forceinline__ double prox_l2_arg(const DProx& p, double x, int64_t i) { return p.shift ? x + p.shift[i] : x; }
// scale = max(0, 1 - lambda * gamma / |z|)
__host__ __device__ inline double prox_l2_scale(double lambda, double gamma, double sumsq) {
  return jl_max(0.0, 1.0 - lambda * gamma / sqrt(sumsq));
}

// contribution of element i to the value f(x) (L1: lambda*|.| summed; L2: squares,
// finished by prox_value_finish; indicators: count of violations)
__device__ __forceinline__ double prox_value_elem(const DProx& p, double x, int64_t i) {
  const double z = p.shift ? x + p.shift[i] : x;
  switch (p.kind) {
    case ADAPROX_P_ZERO: return 0.0;
    case ADAPROX_P_IND_ZERO: return z != 0.0 ? 1.0 : 0.0;
    case ADAPROX_P_NORM_L1: return fabs(z);
    case ADAPROX_P_NORM_L2: return z * z;
    default: {
      const double lo = p.lo_vec ? p.lo_vec[i] : p.lo, hi = p.hi_vec ? p.hi_vec[i] : p.hi;
      return (z < lo || z > hi) ? 1.0 : 0.0;
    }
  }
}
__host__ __device__ inline double prox_value_finish(int kind, double lambda, double total) {
  switch (kind) {
    case ADAPROX_P_ZERO: return 0.0;
    case ADAPROX_P_NORM_L1: return lambda * total;
    case ADAPROX_P_NORM_L2: return lambda * sqrt(total);
    default: return total > 0.0 ? INFINITY : 0.0;
  }
}

// ---------------------------------------------------------------------------
// stepsize rules (src/AdaProx.jl:208-308); state = (s0, s1)
// ---------------------------------------------------------------------------
__host__ __device__ inline void rule_init(const DOpts& o, double& gamma, double& sigma, double& s0, double& s1) {
  gamma = o.gamma;
  switch (o.rule) {
    case ADAPROX_RULE_FIXED: sigma = o.gamma * (o.t * o.t); s0 = 0.0; s1 = 0.0; break;          // :213-215
    case ADAPROX_RULE_MM: sigma = o.gamma * (o.t * o.t); s0 = o.gamma; s1 = INFINITY; break;    // :222-224
    case ADAPROX_RULE_OUR: sigma = o.gamma * (o.t * o.t); s0 = o.gamma; s1 = o.gamma; break;    // :252-256
    default: sigma = o.gamma; s0 = o.gamma; s1 = o.gamma; break;                                // :294-297
  }
}

// AdaPGM is AdaPDM with A = 0 (a scalar), y = zero(x), h = Zero (src/AdaProx.jl:418-421).  Its dual residual (:344-347)
//   w = y + sigma * ((1 + rho) * A_x - rho * A_x_prev),  y+ = prox(IndZero) = 0,  dual_res = (w - y+) / sigma - A_x
// is zero -- unless the new sigma or rho = gamma / gamma_prev is not finite (or sigma = 0): then 0 * Inf / 0 * NaN / 0 / 0 make
// every entry NaN, norm_res (:348) is NaN, the test :354 fails and the next v = x - gamma * grad carries the NaN into x.  The
// kernels without a dual vector reproduce that with the same arithmetic on one entry.  Returns (entry of dual_res)^2.
__host__ __device__ inline double adapgm_dual_res_sq(double gamma, double gamma_prev, double sigma) {
  // ordinary values: rho and 1 + rho are finite, sigma is finite and not 0, every product below is a zero -- skip the two divisions
  // (they sit on the serial scalar path of a 7 us iteration).  NaN fails every comparison and takes the exact path.
  if (fabs(gamma) <= 1e150 && fabs(gamma_prev) >= 1e-150 && fabs(sigma) <= 1.7e308 && sigma != 0.0) return 0.0;
  const double rho = gamma / gamma_prev;                                                       // :342
  const double w = 0.0 + sigma * ((1.0 + rho) * 0.0 - rho * 0.0);                              // :344
  const double d = (w - 0.0) / sigma - 0.0;                                                    // :347
  return d * d;
}

// dgg = sum (dgrad)^2, dgx = <dgrad, dx>, dxx = sum (dx)^2
__host__ __device__ inline void rule_step(const DOpts& o, double dgg, double dgx, double dxx, double& gamma,
                                          double& sigma, double& s0, double& s1) {
  switch (o.rule) {
    case ADAPROX_RULE_FIXED:                                                                   // :213-215
      gamma = o.gamma; sigma = o.gamma * (o.t * o.t);
      return;
    case ADAPROX_RULE_MM: {                                                                    // :226-230
      const double gamma_prev = s0, rho = s1;
      const double L = sqrt(dgg) / sqrt(dxx);
      gamma = jl_min(sqrt(1.0 + rho) * gamma_prev, 1.0 / (2.0 * L));
      sigma = gamma * (o.t * o.t);
      s0 = gamma; s1 = gamma / gamma_prev;
      return;
    }
    case ADAPROX_RULE_OUR: {                                                                   // :258-273
      const double gamma1 = s0, gamma0 = s1;
      const double xi = (o.t * o.t) * (gamma1 * gamma1) * (o.norm_A * o.norm_A);
      const double C = nan_to_zero(norm_sq_jl(dgg) / dgx);
      const double L = nan_to_zero(dgx / norm_sq_jl(dxx));
      const double D = gamma1 * L * (gamma1 * C - 1.0);
      const double d1 = 1.0 + o.delta;
      const double m4 = 1.0 - 4.0 * xi * (d1 * d1);
      gamma = jl_min(jl_min(gamma1 * sqrt(1.0 + gamma1 / gamma0), 1.0 / (2.0 * o.Theta * o.t * o.norm_A)),
                     gamma1 * sqrt(m4) / sqrt(2.0 * d1 * (D + sqrt(D * D + xi * m4))));
      sigma = gamma * (o.t * o.t);
      s0 = gamma; s1 = gamma1;
      return;
    }
    default: {                                                                                 // :299-308
      const double gamma1 = s0, gamma0 = s1;
      const double C = nan_to_zero(norm_sq_jl(dgg) / dgx);
      const double L = nan_to_zero(dgx / norm_sq_jl(dxx));
      const double D = nan_to_zero(1.0 - 2.0 * o.r + gamma1 * L * (gamma1 * C + 2.0 * (o.r - 1.0)));
      gamma = gamma1 * jl_min(sqrt(1.0 / (o.r * (o.nu + o.xi)) + gamma1 / gamma0),
                              sqrt((o.nu * (1.0 + o.xi) - 1.0) / (o.nu * (o.nu + o.xi))) / sqrt(jl_max(D, 0.0)));
      sigma = gamma;
      s0 = gamma; s1 = gamma1;
      return;
    }
  }
}

// ---------------------------------------------------------------------------
// smooth term: eval_with_pullback split into grid phases
//   pre: Gram-form Quadratic only: u = Z'x (phases_pre.cuh; two grid barriers of its own)
//   A: matrix pass F*x           (+ |x|^2 partial for the cubic term; Gram form: Z*u)
//   B: row-wise finalize -> r (what the pullback multiplies), value partials
//   C: matrix pass F'*r          (least squares, logistic)
//   grad_slice: gradient entries of a column slice
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool f_has_gemv_n(int k) {
  return k == ADAPROX_F_LEAST_SQUARES || k == ADAPROX_F_LOGISTIC || k == ADAPROX_F_QUADRATIC || k == ADAPROX_F_CUBIC ||
         k == ADAPROX_F_QUADRATIC_GRAM;
}
__device__ __forceinline__ bool f_has_gemv_t(int k) { return k == ADAPROX_F_LEAST_SQUARES || k == ADAPROX_F_LOGISTIC; }

__device__ __forceinline__ void f_phase_A(const DProblem& P, const DWork& W, const double* x, Sh& sh, double* s_scr,
                                          int b, int G) {
  if (f_has_gemv_n(P.f_kind)) gemv_n_phase(P.F, P.f_kind == ADAPROX_F_QUADRATIC_GRAM ? W.fu : x, sh, b, G);
  if (P.f_kind == ADAPROX_F_CUBIC) {
    double acc[1] = {0.0};
    const int64_t tid = (int64_t)b * kThreads + threadIdx.x, nt = (int64_t)G * kThreads;
    for (int64_t j = tid; j < P.n; j += nt) { const double xv = ldcg(x + j); acc[0] = fma(xv, xv, acc[0]); }
    block_reduce_store<1>(acc, W.red, G, SLOT_XX0, s_scr);
  }
}

__device__ __forceinline__ void f_phase_B(const DProblem& P, const DWork& W, const double* x, double* s_scr, int b, int G) {
  double acc[2] = {0.0, 0.0};
  const int64_t tid = (int64_t)b * kThreads + threadIdx.x, nt = (int64_t)G * kThreads;
  switch (P.f_kind) {
    case ADAPROX_F_LEAST_SQUARES:                                   // lasso/runme.jl:22
      for (int64_t i = tid; i < P.F.m; i += nt) {
        const double res = zsum(P.F, i) - P.fvec[i];
        W.r[i] = res;
        acc[0] = fma(res, res, acc[0]);
      }
      break;
    case ADAPROX_F_LOGISTIC: {                                      // sparse_logreg/runme.jl:24-36
      const double w_end = ldcg(x + (P.n - 1));
      for (int64_t i = tid; i < P.F.m; i += nt) {
        const double logits = zsum(P.F, i) + w_end;
        const double u = 1.0 + exp(-logits);
        const double yi = P.fvec[i];
        const double ri = 1.0 / u - yi;                             // probs - y
        W.r[i] = ri;
        acc[0] += (yi - 1.0) * logits - log(u);
        acc[1] += ri;
      }
    } break;
    case ADAPROX_F_QUADRATIC:                                       // dual_svm/runme.jl:25-27
    case ADAPROX_F_QUADRATIC_GRAM:                                  // same with temp = Z*(Z'x): zsum holds Z*u here
      // (row-sharded Q / Z: this rank owns rows [f_row0, f_row0 + F.m); the sums and the gradient are completed across
      //  the ranks by the caller.  Unsharded: f_row0 = 0 and F.m = n.)
      for (int64_t i = tid; i < P.F.m; i += nt) {
        const int64_t gi = P.f_row0 + i;
        const double temp = zsum(P.F, i), xi = ldcg(x + gi);
        W.r[gi] = temp;
        acc[0] = fma(xi, temp, acc[0]);
        acc[1] = fma(xi, P.fvec[gi], acc[1]);
      }
      break;
    case ADAPROX_F_CUBIC: {                                         // cubic_sparse_logreg/runme.jl:27-29
      double tot[1];
      grid_totals<1>(W.red, G, SLOT_XX0, tot, s_scr);
      const double coef = sqrt(tot[0]) * P.f_c / 2.0;
      for (int64_t i = tid; i < P.n; i += nt) {
        const double xi = ldcg(x + i);
        const double gi = zsum(P.F, i) + P.fvec[i] + coef * xi;
        W.r[i] = gi;
        acc[0] = fma(xi, gi, acc[0]);
        acc[1] = fma(P.fvec[i], xi, acc[1]);
      }
    } break;
    case ADAPROX_F_WORST_QUADRATIC: {                               // nesterov_worst_case/runme.jl:19-39
      const int64_t k = P.f_ipar;
      const double L4 = P.f_c / 4.0;
      for (int64_t i = tid; i < P.n; i += nt) {
        double gi = 0.0;
        if (i < k) {
          const double xi = ldcg(x + i);
          const double xm = (i > 0) ? ldcg(x + i - 1) : 0.0;
          const double xp = (i + 1 < k) ? ldcg(x + i + 1) : 0.0;
          if (i == 0) gi = L4 * (2.0 * xi - xp - 1.0);
          else if (i == k - 1) gi = L4 * (2.0 * xi - xm);
          else gi = L4 * (2.0 * xi - xm - xp);
          if (i + 1 < k) acc[0] += (xi - xp) * (xi - xp);
          if (i == 0) { acc[0] += xi * xi; acc[1] += xi; }
          if (i == k - 1) acc[0] += xi * xi;
        }
        W.r[i] = gi;
      }
    } break;
    case ADAPROX_F_SIMPLE2D:                                        // test/runtests.jl:8-11
      if (tid == 0) {
        const double x1 = ldcg(x), x2 = ldcg(x + 1);
        const double l = log(1.0 + x1 * x1);
        W.r[0] = 2.0 * l * 2.0 * x1 / (1.0 + x1 * x1);
        W.r[1] = 20.0 * x2;
        acc[0] = l * l + 10.0 * (x2 * x2);
      }
      break;
    default:                                                        // ZERO: least_absolute_deviation/runme.jl:18-21
      break;
  }
  block_reduce_store<2>(acc, W.red, G, SLOT_F0, s_scr);
}

// Dense F with a single column chunk (n <= 2048) and a transposed pass (least squares, logistic): both sweeps deal the
// same row blocks to the same CTAs (Segs), so a CTA's A'r sweep consumes exactly the rows of F*x it produced itself.
// Phases A -> B -> C then need CTA barriers only -- two grid barriers less per evaluation, which is what a small
// (latency-bound) problem spends its time on.  B has to visit the CTA's own rows instead of a grid-stride loop.
__device__ __forceinline__ bool f_rows_local(const DProblem& P) {
  return f_has_gemv_t(P.f_kind) && P.F.kind == MAT_DENSE && P.F.nchunks == 1;
}
__device__ __forceinline__ void f_phase_B_local(const DProblem& P, const DWork& W, const double* x, double* s_scr, int b, int G) {
  double acc[2] = {0.0, 0.0};
  Segs sg;
  sg.init(1, P.F.nrb, P.F.rb, P.F.m, b, G);
  const int64_t r0 = sg.T > 0 ? sg.row0 : 0, r1 = sg.T > 0 ? sg.end0 : 0;
  __syncthreads();                                                  // this CTA's zpart rows are complete
  if (P.f_kind == ADAPROX_F_LEAST_SQUARES) {                        // lasso/runme.jl:22
    for (int64_t i = r0 + threadIdx.x; i < r1; i += kThreads) {
      const double res = zsum(P.F, i) - P.fvec[i];
      W.r[i] = res;
      acc[0] = fma(res, res, acc[0]);
    }
  } else {                                                          // logistic, sparse_logreg/runme.jl:24-36
    const double w_end = ldcg(x + (P.n - 1));
    for (int64_t i = r0 + threadIdx.x; i < r1; i += kThreads) {
      const double logits = zsum(P.F, i) + w_end;
      const double u = 1.0 + exp(-logits);
      const double yi = P.fvec[i];
      const double ri = 1.0 / u - yi;
      W.r[i] = ri;
      acc[0] += (yi - 1.0) * logits - log(u);
      acc[1] += ri;
    }
  }
  block_reduce_store<2>(acc, W.red, G, SLOT_F0, s_scr);             // ends with a CTA barrier: r rows visible to phase C
}

__device__ __forceinline__ void f_phase_C(const DProblem& P, const DWork& W, Sh& sh, int b, int G) {
  if (f_has_gemv_t(P.f_kind)) gemv_t_phase(P.F, W.r, sh, b, G);
}

// f(x) from the two value sums
__device__ __forceinline__ double f_value(const DProblem& P, double s0, double s1, double xx) {
  switch (P.f_kind) {
    case ADAPROX_F_LEAST_SQUARES: return 0.5 * norm_sq_jl(s0);              // 0.5 * norm(res)^2
    case ADAPROX_F_LOGISTIC: return -(s0 / P.f_N);                          // -mean(...)
    case ADAPROX_F_QUADRATIC:
    case ADAPROX_F_QUADRATIC_GRAM: return 0.5 * s0 + s1;
    case ADAPROX_F_CUBIC: { const double nx = sqrt(xx); return (s0 + s1) / 2.0 - nx * nx * nx * P.f_c / 12.0; }
    case ADAPROX_F_WORST_QUADRATIC: return (P.f_c / 4.0) * (s0 / 2.0 - s1);
    case ADAPROX_F_SIMPLE2D: return s0;
    default: return 0.0;
  }
}

// gradient entries [j0, j1) -> out; needs the F1 total for the logistic intercept.
__device__ __forceinline__ void grad_slice(const DProblem& P, const DWork& W, int64_t j0, int64_t j1, double* out,
                                           double f1_total, int G) {
  switch (P.f_kind) {
    case ADAPROX_F_LEAST_SQUARES:
      gsum_slice(P.F, j0, j1, out, G);
      break;
    case ADAPROX_F_LOGISTIC: {
      const int64_t je = (j1 < P.n - 1) ? j1 : P.n - 1;
      gsum_slice(P.F, j0, je, out, G);
      for (int64_t j = j0 + threadIdx.x; j < je; j += kThreads) out[j] = out[j] / P.f_N;
      if (j1 == P.n && threadIdx.x == 0) out[P.n - 1] = f1_total / P.f_N;
      __syncthreads();
    } break;
    case ADAPROX_F_QUADRATIC:
    case ADAPROX_F_QUADRATIC_GRAM:
      for (int64_t j = j0 + threadIdx.x; j < j1; j += kThreads) {
        const bool own = (j >= P.f_row0) && (j < P.f_row0 + P.F.m);        // rows of other ranks: 0 (summed across ranks later)
        out[j] = own ? ldcg(W.r + j) + P.fvec[j] : 0.0;
      }
      __syncthreads();
      break;
    case ADAPROX_F_ZERO:
      for (int64_t j = j0 + threadIdx.x; j < j1; j += kThreads) out[j] = 0.0;
      __syncthreads();
      break;
    default:
      for (int64_t j = j0 + threadIdx.x; j < j1; j += kThreads) out[j] = ldcg(W.r + j);
      __syncthreads();
      break;
  }
}

}  // namespace adaprox
