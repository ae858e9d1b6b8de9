// solver_pg.cuh -- the proximal-gradient baselines AdaPGM is compared against,
// as one persistent cooperative kernel:
//   backtracking_proxgrad   src/AdaProx.jl:34-64
//   backtracking_nesterov   src/AdaProx.jl:66-84
//   fixed_nesterov          src/AdaProx.jl:91-142
//   agraal                  src/AdaProx.jl:150-192
// Data-dependent control flow (the `while f_z > ub_z` backtracking loop) runs on
// the device: every CTA derives the same scalars from fixed-order reductions and
// takes the same branch, so there is no host round trip per trial.
#pragma once
#include "phases_pre.cuh"

namespace adaprox {

// f(x) (and optionally the gradient on this CTA's slice).  `x` must be complete
// and grid-synced.  Returns with all partials consumed; the caller must
// grid-sync before anything overwrites W.r / the matrix partials again.
// Row-sharded f (least squares: rows of A and b local; Quadratic: rows of Q local): the value sums and the gradient are
// completed across the ranks inside the kernel (p2p.cuh); every rank then continues with identical bits.
__device__ __forceinline__ double eval_f_grid(cg::grid_group& grid, const DProblem& P, const DWork& W, const double* x,
                                              bool want_grad, double* grad_out, int64_t j0, int64_t j1, Sh& sh,
                                              double* s_scr, int b, int G, P2PState& ps) {
  const bool shardedF = P.p2p.n > 1 && P.F_sharded;
  f_phase_pre(grid, P, W, x, sh, b, G, &ps);
  if (f_rows_local(P)) {          // phases A -> B -> C on this CTA's own rows: CTA barriers only
    f_phase_A(P, W, x, sh, s_scr, b, G);
    f_phase_B_local(P, W, x, s_scr, b, G);
  } else {
    f_phase_A(P, W, x, sh, s_scr, b, G);
    grid.sync();
    f_phase_B(P, W, x, s_scr, b, G);
  }
  if (!(want_grad && f_rows_local(P))) grid.sync();
  if (want_grad) {
    f_phase_C(P, W, sh, b, G);
    grid.sync();
  }
  double tot[2], xx[1] = {0.0};
  grid_totals<2>(W.red, G, SLOT_F0, tot, s_scr);
  if (shardedF) p2p_allreduce_scalars<2>(P.p2p, ps, grid, tot);
  if (P.f_kind == ADAPROX_F_CUBIC) grid_totals<1>(W.red, G, SLOT_XX0, xx, s_scr);
  if (want_grad) {
    grad_slice(P, W, j0, j1, grad_out, tot[1], G);
    if (shardedF) p2p_allreduce<kThreads>(P.p2p, ps, grid, grad_out, grad_out, P.n);
  }
  return f_value(P, tot[0], tot[1], xx[0]);
}

__global__ void __launch_bounds__(kThreads, 2) k_proxgrad_family(DProblem P, DOpts O, DWork W) {
  cg::grid_group grid = cg::this_grid();
  const int b = blockIdx.x, G = gridDim.x;
  extern __shared__ __align__(1024) unsigned char dyn_smem[];
  __shared__ double s_scr[kWarps * 8 + kMaxRed];
  __shared__ double s_part[kPartRows * kWarps];
  __shared__ unsigned long long s_bars[2 * kStages];
  Sh sh;
  sh_init(sh, dyn_smem, s_scr, s_part, s_bars);
  P2PState ps;
  p2p_begin(P.p2p, ps);
  const int64_t tid = (int64_t)b * kThreads + threadIdx.x, nt = (int64_t)G * kThreads;
  int64_t j0, j1;
  cta_slice(P.n, b, G, j0, j1);
  const bool want_obj = O.want_objective != 0;

  int64_t n_eval = 0, n_grad = 0, n_proxg = 0, n_rec = 0;
  unsigned flags = 0;
  double gamma = O.gamma, norm_res = INFINITY;
  int64_t it_done = O.maxit;
  bool converged = false;
  double* result = W.xb[0];

  auto record = [&](int64_t it, double objective_f, double objective_g) {
    if (b == 0 && threadIdx.x == 0 && W.rec != nullptr && it <= O.max_records) {
      adaprox_record rc;
      rc.it = it; rc.gamma = gamma; rc.sigma = NAN; rc.norm_res = norm_res;
      rc.f_x = objective_f; rc.g_x = objective_g; rc.h_Ax = 0.0;
      rc.f_evals = n_eval; rc.grad_f_evals = n_grad; rc.prox_g_evals = n_proxg; rc.prox_h_evals = 0;
      rc.A_evals = 0; rc.At_evals = 0;
      W.rec[it - 1] = rc;
    }
    if (it <= O.max_records) n_rec = it;
  };

  if (O.solver == ADAPROX_S_BACKTRACKING_PROXGRAD || O.solver == ADAPROX_S_BACKTRACKING_NESTEROV) {
    const bool nesterov = (O.solver == ADAPROX_S_BACKTRACKING_NESTEROV);
    double* x = W.xb[0];            // x0
    double* z = W.xb[1];
    double* z_prev = W.xb[2];       // Nesterov only (z = x0 initially, :67)
    double* grad = W.gb[0];
    if (nesterov) for (int64_t j = j0 + threadIdx.x; j < j1; j += kThreads) z_prev[j] = x[j];
    double theta = 1.0;                                                          // :68
    double f_x = eval_f_grid(grid, P, W, x, true, grad, j0, j1, sh, s_scr, b, G, ps);   // :52 / :69
    n_eval++; n_grad++;
    int64_t trial_no = 0;
    for (int64_t it = 1; it <= O.maxit; ++it) {
      if (p2p_failed(P.p2p)) { flags |= ADAPROX_FLAG_COMM; break; }
      // ---- backtrack_stepsize (:34-48) ---------------------------------------------
      gamma = nesterov ? gamma : O.xi * gamma;                                   // :54 / :72
      double f_z = 0.0, g_z = 0.0, dzz = 0.0;
      for (;;) {
        const int base = (trial_no & 1) ? SLOT_DR : SLOT_PR;
        ++trial_no;
        double acc[3] = {0.0, 0.0, 0.0};
        for (int64_t j = j0 + threadIdx.x; j < j1; j += kThreads) {
          const double xj = x[j], gj = grad[j];
          const double zj = prox_elem(P.g, xj - gamma * gj, gamma, j, 0.0);      // :35 / :43
          z[j] = zj;
          const double d = zj - xj;
          acc[0] = fma(gj, d, acc[0]);
          acc[1] = fma(d, d, acc[1]);
          acc[2] += prox_value_elem(P.g, zj, j);
        }
        block_reduce_store<3>(acc, W.red, G, base, s_scr);
        n_proxg++;
        grid.sync();
        f_z = eval_f_grid(grid, P, W, z, false, nullptr, j0, j1, sh, s_scr, b, G, ps);   // :37 / :45
        n_eval++;
        double t3[3];
        grid_totals<3>(W.red, G, base, t3, s_scr);
        dzz = t3[1];
        g_z = prox_value_finish(P.g.kind, P.g.lambda, t3[2]);
        const double ub_z = f_x + t3[0] + 1.0 / (2.0 * gamma) * norm_sq_jl(dzz);   // :26
        if (!(f_z > ub_z)) break;                                                // :38
        gamma *= O.shrink;                                                       // :39
        if (gamma < 1e-12) flags |= ADAPROX_FLAG_STEP_TOO_SMALL;                 // :40-42 (the reference keeps looping)
        if (gamma < 1e-300) break;
      }
      norm_res = sqrt(dzz) / gamma;                                              // :55 / :73
      record(it, f_z, g_z);
      result = z;
      if (norm_res <= O.tol) { converged = true; it_done = it; break; }          // :57 / :75
      if (!nesterov) {
        // x, f_x = z, f_z ; grad_x = pb()  (:60-61) -- W.r still holds the residual of z
        f_phase_C(P, W, sh, b, G);
        grid.sync();
        double tot[2];
        grid_totals<2>(W.red, G, SLOT_F0, tot, s_scr);
        grad_slice(P, W, j0, j1, grad, tot[1], G);
        if (P.p2p.n > 1 && P.F_sharded) p2p_allreduce<kThreads>(P.p2p, ps, grid, grad, grad, P.n);
        n_grad++;
        double* t = x; x = z; z = t;
        f_x = f_z;
        result = x;
      } else {
        const double theta_prev = theta;                                         // :78-80
        theta = (1.0 + sqrt(1.0 + 4.0 * theta_prev * theta_prev)) / 2.0;
        const double beta = (theta_prev - 1.0) / theta;
        for (int64_t j = j0 + threadIdx.x; j < j1; j += kThreads) {
          const double zj = z[j];
          x[j] = zj + beta * (zj - z_prev[j]);
        }
        double* t = z_prev; z_prev = z; z = t;                                   // :71
        result = z_prev;
        grid.sync();
        f_x = eval_f_grid(grid, P, W, x, true, grad, j0, j1, sh, s_scr, b, G, ps);  // :81
        n_eval++; n_grad++;
      }
    }
  } else if (O.solver == ADAPROX_S_FIXED_NESTEROV) {
    const double mu = O.muf + O.mug;                                             // :108-117
    const double q = gamma * mu / (1.0 + gamma * O.mug);
    double theta = O.theta >= 0.0 ? O.theta : (q > 0.0 ? 1.0 / sqrt(q) : 0.0);
    double* x = W.xb[0];
    double* x_prev = W.xb[1];
    double* z = W.xb[2];
    double* grad = W.gb[0];
    for (int64_t j = j0 + threadIdx.x; j < j1; j += kThreads) x_prev[j] = x[j];  // :119
    for (int64_t it = 1; it <= O.maxit; ++it) {
      if (p2p_failed(P.p2p)) { flags |= ADAPROX_FLAG_COMM; break; }
      const double theta_prev = theta;
      double beta;
      if (mu == 0.0) {                                                           // :122-128
        theta = (1.0 + sqrt(1.0 + 4.0 * theta_prev * theta_prev)) / 2.0;
        beta = (theta_prev - 1.0) / theta;
      } else {
        const double a = 1.0 - q * theta_prev * theta_prev;
        theta = (a + sqrt(a * a + 4.0 * theta_prev * theta_prev)) / 2.0;
        beta = (theta_prev - 1.0) * (1.0 + gamma * O.mug - theta * gamma * mu) / theta / (1.0 - gamma * O.muf);
      }
      for (int64_t j = j0 + threadIdx.x; j < j1; j += kThreads) z[j] = x[j] + beta * (x[j] - x_prev[j]);   // :129
      grid.sync();
      eval_f_grid(grid, P, W, z, true, grad, j0, j1, sh, s_scr, b, G, ps);          // :130
      n_eval++; n_grad++;
      double acc[2] = {0.0, 0.0};
      for (int64_t j = j0 + threadIdx.x; j < j1; j += kThreads) {               // :131-133 (x_prev <- x, x <- prox)
        const double zj = z[j];
        const double xn = prox_elem(P.g, zj - gamma * grad[j], gamma, j, 0.0);
        x_prev[j] = xn;                       // written into the dead buffer; pointers swap below
        const double d = xn - zj;
        acc[0] = fma(d, d, acc[0]);
        acc[1] += prox_value_elem(P.g, xn, j);
      }
      const int base = (it & 1) ? SLOT_DR : SLOT_PR;
      block_reduce_store<2>(acc, W.red, G, base, s_scr);
      n_proxg++;
      { double* t = x; x = x_prev; x_prev = t; }
      grid.sync();
      double t2[2];
      grid_totals<2>(W.red, G, base, t2, s_scr);
      norm_res = sqrt(t2[0]) / gamma;
      double fx = NAN;
      if (want_obj) {                                                            // :134-136, uncounted f(x)
        fx = eval_f_grid(grid, P, W, x, false, nullptr, j0, j1, sh, s_scr, b, G, ps);
        grid.sync();
      }
      record(it, fx, prox_value_finish(P.g.kind, P.g.lambda, t2[1]));
      result = x;
      if (norm_res <= O.tol) { converged = true; it_done = it; break; }
    }
  } else {   // ADAPROX_S_AGRAAL (:150-192); W.yb[0] holds the second start point x0
    double* x = W.xb[0];
    double* x_prev = W.xb[1];
    double* x_bar = W.xb[2];
    double* grad = W.gb[0];
    double* grad_prev = W.gb[1];
    for (int64_t j = j0 + threadIdx.x; j < j1; j += kThreads) { x_prev[j] = W.aux[0][j]; x_bar[j] = x[j]; }   // :165
    grid.sync();
    eval_f_grid(grid, P, W, x, true, grad, j0, j1, sh, s_scr, b, G, ps);            // :166
    grid.sync();
    eval_f_grid(grid, P, W, x_prev, true, grad_prev, j0, j1, sh, s_scr, b, G, ps);  // :167
    n_eval = 2; n_grad = 2;
    const double phi = O.phi;
    const double rho = 1.0 / phi + 1.0 / (phi * phi);                            // :172
    double theta = 1.0;
    for (int64_t it = 1; it <= O.maxit; ++it) {
      if (p2p_failed(P.p2p)) { flags |= ADAPROX_FLAG_COMM; break; }
      const int base = (it & 1) ? SLOT_DR : SLOT_PR;
      double acc[2] = {0.0, 0.0};
      for (int64_t j = j0 + threadIdx.x; j < j1; j += kThreads) {
        const double dx = x[j] - x_prev[j], dg = grad[j] - grad_prev[j];
        acc[0] = fma(dx, dx, acc[0]);
        acc[1] = fma(dg, dg, acc[1]);
      }
      block_reduce_store<2>(acc, W.red, G, base, s_scr);
      grid.sync();
      double t2[2];
      grid_totals<2>(W.red, G, base, t2, s_scr);
      if (it == 1 && !(O.gamma > 0.0)) gamma = sqrt(t2[0]) / sqrt(t2[1]);        // :168-170
      const double C = norm_sq_jl(t2[0]) / norm_sq_jl(t2[1]);                    // :175
      const double gamma_prev = gamma;
      gamma = jl_min(jl_min(rho * gamma_prev, phi * theta * C / (4.0 * gamma_prev)), O.gamma_max);   // :177
      theta = phi * gamma / gamma_prev;                                          // :178
      double acc2[2] = {0.0, 0.0};
      for (int64_t j = j0 + threadIdx.x; j < j1; j += kThreads) {
        const double xj = x[j];
        const double xb = ((phi - 1.0) * xj + x_bar[j]) / phi;                   // :179
        x_bar[j] = xb;
        const double xn = prox_elem(P.g, xb - gamma * grad[j], gamma, j, 0.0);   // :181
        x_prev[j] = xn;                       // dead buffer; swapped below (:180)
        const double d = xn - xj;
        acc2[0] = fma(d, d, acc2[0]);
        acc2[1] += prox_value_elem(P.g, xn, j);
      }
      block_reduce_store<2>(acc2, W.red, G, base + 2, s_scr);
      n_proxg++;
      { double* t = x; x = x_prev; x_prev = t; }
      { double* t = grad; grad = grad_prev; grad_prev = t; }
      grid.sync();
      double t3[2];
      grid_totals<2>(W.red, G, base + 2, t3, s_scr);
      norm_res = sqrt(t3[0]) / gamma;                                            // :182
      double fx = NAN;
      if (want_obj) {
        fx = eval_f_grid(grid, P, W, x, false, nullptr, j0, j1, sh, s_scr, b, G, ps);
        grid.sync();
      }
      record(it, fx, prox_value_finish(P.g.kind, P.g.lambda, t3[1]));
      result = x;
      if (norm_res <= O.tol) { converged = true; it_done = it; break; }
      eval_f_grid(grid, P, W, x, true, grad, j0, j1, sh, s_scr, b, G, ps);          // :189
      n_eval++; n_grad++;
    }
  }

  if (!(gamma == gamma) || !(norm_res == norm_res)) flags |= ADAPROX_FLAG_NONFINITE;
  grid.sync();
  for (int64_t j = tid; j < P.n; j += nt) W.xout[j] = ldcg(result + j);
  if (b == 0 && threadIdx.x == 0) {
    DResult r;
    r.iters = it_done;
    r.flags = flags | (converged ? ADAPROX_FLAG_CONVERGED : 0u);
    r.xbuf = 0;
    r.f_evals = n_eval; r.grad_f_evals = n_grad; r.prox_g_evals = n_proxg; r.prox_h_evals = 0;
    r.A_evals = 0; r.At_evals = 0; r.n_records = n_rec;
    r.final_gamma = gamma; r.final_sigma = NAN; r.final_norm_res = norm_res;
    *W.res = r;
  }
}

}  // namespace adaprox
