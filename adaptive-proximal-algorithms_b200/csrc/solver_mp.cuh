// solver_mp.cuh -- Malitsky-Pock primal-dual linesearch (src/AdaProx.jl:555-629) as a
// persistent cooperative kernel.  The PD competitor of every AdaPDM figure
// (experiments/dual_svm/runme.jl:78-92, least_absolute_deviation/runme.jl:64-78).
//
// One deliberate economy: the reference evaluates f and its gradient at the new x at the
// end of an iteration (`grad_x = pb()`, :612) and again at the top of the next one
// (`eval_with_gradient(f, x)`, :608) -- the same point, bit-identical values.  The kernel
// evaluates once and counts twice, so the Counting figures match the reference exactly.
#pragma once
#include "phases_pre.cuh"

namespace adaprox {

__global__ void __launch_bounds__(kThreads, 2) k_malitsky_pock(DProblem P, DOpts O, DWork W) {
  cg::grid_group grid = cg::this_grid();
  const int b = blockIdx.x, G = gridDim.x;
  extern __shared__ __align__(1024) unsigned char dyn_smem[];
  __shared__ double s_scr[kWarps * 8 + kMaxRed];
  __shared__ double s_part[kPartRows * kWarps];
  __shared__ unsigned long long s_bars[2 * kStages];
  Sh sh;
  sh_init(sh, dyn_smem, s_scr, s_part, s_bars);
  const int64_t tid = (int64_t)b * kThreads + threadIdx.x, nt = (int64_t)G * kThreads;
  int64_t j0, j1;
  cta_slice(P.n, b, G, j0, j1);
  const bool want_obj = O.want_objective != 0;
  const bool h_l2 = (P.h.kind == ADAPROX_P_NORM_L2);

  double sigma = O.sigma, gamma = 0.0, norm_res = INFINITY;
  const double t = O.t, theta = 1.0;                              // :595 (never updated by the reference)
  int64_t n_eval = 0, n_grad = 0, n_proxg = 0, n_proxh = 0, n_mul = 0, n_amul = 0, n_rec = 0;
  unsigned flags = 0;
  int64_t it_done = O.maxit;
  bool converged = false;

  double* x_prev = W.xb[0];        // holds x0
  double* x = W.xb[1];
  double* y = W.yb[0];             // holds y0 (updated in place: y_prev is never read, :596,614)
  double* Ax_prev = W.Axb[0];
  double* Ax = W.Axb[1];
  double* grad_prev = W.gb[0];
  double* grad = W.gb[1];
  int atc = 0;

  // ---- prologue: A_x = A*x, At_y = A'*y (:597-598) and f, grad at x0 (first :608) ------------------
  f_phase_pre(grid, P, W, x_prev, sh, b, G, nullptr);
  gemv_n_phase(P.A, x_prev, sh, b, G);
  f_phase_A(P, W, x_prev, sh, s_scr, b, G);
  grid.sync();
  for (int64_t i = tid; i < P.md; i += nt) Ax_prev[i] = zsum(P.A, i);
  f_phase_B(P, W, x_prev, s_scr, b, G);
  grid.sync();
  gemv_t_phase(P.A, y, sh, b, G);
  f_phase_C(P, W, sh, b, G);
  grid.sync();
  double ftot[2], xx0[1] = {0.0};
  grid_totals<2>(W.red, G, SLOT_F0, ftot, s_scr);
  if (P.f_kind == ADAPROX_F_CUBIC) grid_totals<1>(W.red, G, SLOT_XX0, xx0, s_scr);
  double f_x_prev = f_value(P, ftot[0], ftot[1], xx0[0]);
  grad_slice(P, W, j0, j1, grad_prev, ftot[1], G);
  gsum_slice(P.A, j0, j1, W.Aty[atc], G);
  n_mul = 1; n_amul = 1;
  grid.sync();
  int64_t trial_no = 0;

  for (int64_t it = 1; it <= O.maxit; ++it) {
    // ---- dual step: w = y + sigma*A_x ; y = prox(h*, w, sigma)  (:601-602) --------------------------
    double l2acc[1] = {0.0};
    for (int64_t i = tid; i < P.md; i += nt) {
      const double wi = y[i] + sigma * Ax_prev[i];
      W.w[i] = wi;
      if (h_l2) { const double z = prox_l2_arg(P.h, wi / sigma, i); l2acc[0] = fma(z, z, l2acc[0]); }
    }
    double l2scale = 0.0;
    if (h_l2) {
      block_reduce_store<1>(l2acc, W.red, G, SLOT_L2, s_scr);
      grid.sync();
      double l2tot[1];
      grid_totals<1>(W.red, G, SLOT_L2, l2tot, s_scr);
      l2scale = prox_l2_scale(P.h.lambda, 1.0 / sigma, l2tot[0]);
    }
    for (int64_t i = tid; i < P.md; i += nt) y[i] = prox_conj_elem(P.h, W.w[i], sigma, i, l2scale);
    n_proxh++;
    grid.sync();
    gemv_t_phase(P.A, y, sh, b, G);                               // At_y = A'*y (:603)
    n_amul++;
    grid.sync();
    const double* Aty_prev = W.Aty[atc];
    double* Aty = W.Aty[atc ^ 1];
    gsum_slice(P.A, j0, j1, Aty, G);
    atc ^= 1;
    const double sigma_prev = sigma;                              // :605-606
    sigma = sigma * sqrt(1.0 + theta);
    n_eval++; n_grad++;                                           // :608 (same point as the last evaluation: reused)

    // ---- backtrack_stepsize_MP (:555-579) -------------------------------------------------------------
    double f_x = 0.0, gval = 0.0;
    for (;;) {
      const int base = (trial_no & 1) ? SLOT_DR : SLOT_PR;
      ++trial_no;
      const double th = sigma / sigma_prev;                       // :556 / :569
      gamma = t * t * sigma;                                      // :557 / :570
      double acc[3] = {0.0, 0.0, 0.0};
      for (int64_t j = j0 + threadIdx.x; j < j1; j += kThreads) {
        const double aty_bar = (1.0 + th) * Aty[j] - th * Aty_prev[j];           // :558
        const double vj = x_prev[j] - gamma * (aty_bar + grad_prev[j]);          // :559
        W.v[j] = vj;
        const double xj = prox_elem(P.g, vj, gamma, j, 0.0);                     // :560
        x[j] = xj;
        const double d = xj - x_prev[j];
        acc[0] = fma(d, d, acc[0]);
        acc[1] = fma(grad_prev[j], d, acc[1]);
        if (want_obj) acc[2] += prox_value_elem(P.g, xj, j);
      }
      block_reduce_store<3>(acc, W.red, G, base, s_scr);
      n_proxg++;
      grid.sync();
      f_phase_pre(grid, P, W, x, sh, b, G, nullptr);
      gemv_n_phase(P.A, x, sh, b, G);                             // :561
      f_phase_A(P, W, x, sh, s_scr, b, G);                        // :562
      n_mul++; n_eval++;
      grid.sync();
      double dacc[1] = {0.0};
      for (int64_t i = tid; i < P.md; i += nt) {
        const double axi = zsum(P.A, i);
        Ax[i] = axi;
        const double d = axi - Ax_prev[i];
        dacc[0] = fma(d, d, dacc[0]);
      }
      block_reduce_store<1>(dacc, W.red, G, base + 3, s_scr);
      f_phase_B(P, W, x, s_scr, b, G);
      grid.sync();
      double t4[4];
      grid_totals<4>(W.red, G, base, t4, s_scr);                  // |x - x_prev|^2, <grad_prev, dx>, g(x), |A_x - A_x_prev|^2
      grid_totals<2>(W.red, G, SLOT_F0, ftot, s_scr);
      if (P.f_kind == ADAPROX_F_CUBIC) grid_totals<1>(W.red, G, SLOT_XX0, xx0, s_scr);
      f_x = f_value(P, ftot[0], ftot[1], xx0[0]);
      gval = t4[2];
      const double lhs = gamma * sigma * norm_sq_jl(t4[3]) + 2.0 * gamma * (f_x - f_x_prev - t4[1]);   // :563
      if (!(lhs > 0.95 * norm_sq_jl(t4[0]))) break;               // :564
      sigma /= 2.0;                                               // :565
      if (sigma < 1e-12) flags |= ADAPROX_FLAG_STEP_TOO_SMALL;    // :566-568
      if (sigma < 1e-300) break;
    }
    // ---- grad_x = pb() (:612), residuals (:616-618) -----------------------------------------------------
    f_phase_C(P, W, sh, b, G);
    n_grad++;
    grid.sync();
    grad_slice(P, W, j0, j1, grad, ftot[1], G);
    {
      double acc[3] = {0.0, 0.0, 0.0};
      for (int64_t j = j0 + threadIdx.x; j < j1; j += kThreads) {
        const double pr = (W.v[j] - x[j]) / gamma + grad[j] + Aty[j];            // :616
        acc[0] = fma(pr, pr, acc[0]);
      }
      for (int64_t i = tid; i < P.md; i += nt) {
        const double axi = Ax[i];
        const double dr = (W.w[i] - y[i]) / sigma_prev - axi;                    // :617
        acc[1] = fma(dr, dr, acc[1]);
        if (want_obj) acc[2] += prox_value_elem(P.h, axi, i);
      }
      block_reduce_store<3>(acc, W.red, G, SLOT_DY, s_scr);       // slots 11, 12, 13
    }
    grid.sync();
    double t3[3];
    grid_totals<3>(W.red, G, SLOT_DY, t3, s_scr);
    norm_res = sqrt(norm_sq_jl(t3[0]) + norm_sq_jl(t3[1]));      // :618
    if (!(gamma == gamma) || !(norm_res == norm_res)) flags |= ADAPROX_FLAG_NONFINITE;
    if (b == 0 && threadIdx.x == 0 && W.rec != nullptr && it <= O.max_records) {
      adaprox_record rc;
      rc.it = it; rc.gamma = gamma; rc.sigma = sigma; rc.norm_res = norm_res;
      rc.f_x = f_x;
      rc.g_x = want_obj ? prox_value_finish(P.g.kind, P.g.lambda, gval) : NAN;
      rc.h_Ax = want_obj ? prox_value_finish(P.h.kind, P.h.lambda, t3[2]) : NAN;
      rc.f_evals = n_eval; rc.grad_f_evals = n_grad; rc.prox_g_evals = n_proxg; rc.prox_h_evals = n_proxh;
      rc.A_evals = n_mul; rc.At_evals = n_amul;
      W.rec[it - 1] = rc;
    }
    if (it <= O.max_records) n_rec = it;
    { double* tp = x_prev; x_prev = x; x = tp; }                  // the new iterate becomes x_prev of the next iteration
    { double* tp = Ax_prev; Ax_prev = Ax; Ax = tp; }
    { double* tp = grad_prev; grad_prev = grad; grad = tp; }
    f_x_prev = f_x;
    if (norm_res <= O.tol) { converged = true; it_done = it; break; }            // :624-626
  }

  grid.sync();
  for (int64_t j = tid; j < P.n; j += nt) W.xout[j] = ldcg(x_prev + j);
  if (W.yout) for (int64_t i = tid; i < P.md; i += nt) W.yout[i] = ldcg(y + i);
  if (b == 0 && threadIdx.x == 0) {
    DResult r;
    r.iters = it_done;
    r.flags = flags | (converged ? ADAPROX_FLAG_CONVERGED : 0u);
    r.xbuf = 0;
    r.f_evals = n_eval; r.grad_f_evals = n_grad; r.prox_g_evals = n_proxg; r.prox_h_evals = n_proxh;
    r.A_evals = n_mul; r.At_evals = n_amul; r.n_records = n_rec;
    r.final_gamma = gamma; r.final_sigma = sigma; r.final_norm_res = norm_res;
    *W.res = r;
  }
}

}  // namespace adaprox
