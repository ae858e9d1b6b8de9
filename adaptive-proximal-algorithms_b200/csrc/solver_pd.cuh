// solver_pd.cuh -- the generic adaptive primal-dual loop as ONE persistent
// cooperative kernel (src/AdaProx.jl:312-364), plus its linesearch variant
// AdaPDM+ (src/AdaProx.jl:463-550).
//
// AdaPGM (:418-421), fixed-step PGM (:457-459) and Condat-Vu (:367-416) are
// this kernel with A = none / a fixed rule, exactly as in the reference.
// One grid of (SMs x 2) CTAs stays resident for the whole solve; phases are
// separated by grid syncs; every scalar of the stepsize rule is recomputed
// redundantly and bit-identically by every CTA from fixed-order reductions, so
// there is no host round trip and no single-thread serial section.
//
// Per iteration (reference line numbers on the right):
//   P1  F*x partials, A*x partials                               :335-336
//   P2  rows: residual r, f(x) sums, A_x                         :336
//   P3  F'*r partials                                            :336 (pullback)
//   P4  slice: grad, primal_res, |dgrad|^2 <dgrad,dx> |dx|^2     :338, :260-261
//   P5  stepsize (all threads); rows: w, y+ = prox_{sigma h*}    :341-347
//   P6  norm_res, record, convergence; A'*y+ partials            :348-358
//   P7  slice: A'y, v, x+ = prox_{gamma g}(v)                    :359-361
#pragma once
#include "phases_pre.cuh"

namespace adaprox {

// g(x) partials are produced in P7 and consumed in P5 of the next iteration with
// no grid sync between P5 and P7 of the same iteration (AdaPGM): alternate slots.
__device__ __forceinline__ int gval_slot(int64_t it) { return (it & 1) ? SLOT_GVAL : SLOT_AUX0; }

// dual step on this CTA's rows.  Returns via reductions: |dual_res|^2, h(A_x) sum.
// For the NormL2 prox the caller has already reduced |w/sigma + shift|^2 (l2sum).
__device__ __forceinline__ void dual_rows(const DProblem& P, const DWork& W, const double* w, const double* Ax,
                                          double* ynew, double sigma, double l2sum, bool want_h, const double* yold,
                                          double* s_scr, int b, int G, bool linesearch) {
  double acc[3] = {0.0, 0.0, 0.0};
  // h passed as a conjugate (h = phi*): convex_conjugate(h) = phi, so the dual step is the plain prox of the base function
  const bool h_conj = P.h.conjugate != 0;
  const double l2scale = (P.h.kind == ADAPROX_P_NORM_L2) ? prox_l2_scale(P.h.lambda, h_conj ? sigma : 1.0 / sigma, l2sum) : 0.0;
  const int64_t tid = (int64_t)b * kThreads + threadIdx.x, nt = (int64_t)G * kThreads;
  for (int64_t i = tid; i < P.md; i += nt) {
    const double wi = ldcg(w + i), axi = ldcg(Ax + i);
    const double yi = h_conj ? prox_elem(P.h, wi, sigma, i, l2scale) : prox_conj_elem(P.h, wi, sigma, i, l2scale);   // :345
    ynew[i] = yi;
    const double dr = (wi - yi) / sigma - axi;                                   // :347
    acc[0] = fma(dr, dr, acc[0]);
    if (want_h && !h_conj) acc[1] += prox_value_elem(P.h, axi, i);
    if (linesearch) { const double dy = yi - ldcg(yold + i); acc[2] = fma(dy, dy, acc[2]); }
  }
  double a2[2] = {acc[0], acc[1]};
  block_reduce_store<2>(a2, W.red, G, SLOT_DR, s_scr);
  if (linesearch) { double a1[1] = {acc[2]}; block_reduce_store<1>(a1, W.red, G, SLOT_DY, s_scr); }
}

template <bool LINESEARCH>
__global__ void __launch_bounds__(kThreads, 2) k_primal_dual(DProblem P, DOpts O, DWork W) {
  cg::grid_group grid = cg::this_grid();
  const int b = blockIdx.x, G = gridDim.x;
  extern __shared__ __align__(1024) unsigned char dyn_smem[];
  __shared__ double s_scr[kWarps * 8 + kMaxRed];
  __shared__ double s_part[kPartRows * kWarps];
  __shared__ unsigned long long s_bars[2 * kStages];
  Sh sh;
  sh_init(sh, dyn_smem, s_scr, s_part, s_bars);

  const bool hasA = (P.A.kind != MAT_NONE);
  const bool h_l2 = hasA && (P.h.kind == ADAPROX_P_NORM_L2);
  // Row-sharded A: this rank holds md = local rows of A, the dual iterate y and A*x are row-local, x and every
  // n-vector are replicated.  A'y and the dual-side sums are combined across the ranks INSIDE this kernel (p2p.cuh); every
  // rank then holds identical bits and the replicated control flow stays in lock step.
  const bool shardedA = hasA && P.p2p.n > 1 && P.A_sharded;
  // Row-sharded Q of a Quadratic smooth term (dual SVM, dual_svm/runme.jl:19-28): the gradient rows and the two value
  // sums of this rank are completed across the ranks in the same way.
  const bool shardedF = P.p2p.n > 1 && P.F_sharded;
  P2PState ps;
  p2p_begin(P.p2p, ps);
  const bool want_obj = O.want_objective != 0;
  const int64_t tid = (int64_t)b * kThreads + threadIdx.x, nt = (int64_t)G * kThreads;
  int64_t j0, j1;
  cta_slice(P.n, b, G, j0, j1);

  // ---- primal step  v = x - gamma (grad + A'y);  x+ = prox_{gamma g}(v)  (:330-332 / :359-361) ----------------------
  // Separable g: one pass.  g = NormL2 (block soft threshold) needs |v + shift| first, a conjugate g
  // (ProximalCore's Moreau order) |v / gamma + shift|: one more reduction and grid barrier, only for those.
  const bool g_conj = P.g.conjugate != 0;
  const bool g_l2 = (P.g.kind == ADAPROX_P_NORM_L2);
  auto primal_step = [&](const double* xin, const double* grad_in, const double* aty_in, double gam, double* xn, int gslot) {
    double l2scale = 0.0;
    if (g_l2) {
      double a1[1] = {0.0};
      for (int64_t j = j0 + threadIdx.x; j < j1; j += kThreads) {
        const double vj = xin[j] - gam * (grad_in[j] + (aty_in ? aty_in[j] : 0.0));
        const double z = prox_l2_arg(P.g, g_conj ? vj / gam : vj, j);
        a1[0] = fma(z, z, a1[0]);
      }
      block_reduce_store<1>(a1, W.red, G, SLOT_AUX1, s_scr);
      grid.sync();
      double tot[1];
      grid_totals<1>(W.red, G, SLOT_AUX1, tot, s_scr);
      l2scale = prox_l2_scale(P.g.lambda, g_conj ? 1.0 / gam : gam, tot[0]);
    }
    double acc[1] = {0.0};
    for (int64_t j = j0 + threadIdx.x; j < j1; j += kThreads) {
      const double vj = xin[j] - gam * (grad_in[j] + (aty_in ? aty_in[j] : 0.0));
      W.v[j] = vj;
      const double xj = g_conj ? prox_conj_elem(P.g, vj, gam, j, l2scale) : prox_elem(P.g, vj, gam, j, l2scale);
      xn[j] = xj;
      if (want_obj && !g_conj) acc[0] += prox_value_elem(P.g, xj, j);
    }
    block_reduce_store<1>(acc, W.red, G, gslot, s_scr);
  };

  // ---- rule initialisation (:324 / :484-491) ------------------------------
  double gamma, sigma, s0, s1;
  double eta = O.eta, gamma_prev_ls = 0.0;
  if (LINESEARCH) {
    gamma = O.gamma; sigma = O.t * O.t * gamma; s0 = s1 = 0.0;
    gamma_prev_ls = gamma;                                                       // :491
  } else {
    rule_init(O, gamma, sigma, s0, s1);
  }

  int xc = 0, gc = 0, axc = 0, yc = 0, atc = 0;       // ring positions
  double* x = W.xb[0];                                // holds x0
  double* y = W.yb[0];                                // holds y0
  int64_t n_eval = 0, n_grad = 0, n_proxg = 0, n_proxh = 0, n_mul = 0, n_amul = 0, n_rec = 0;
  unsigned flags = 0;

  // ---- prologue (:327-332) --------------------------------------------------
  // rows_local: phases A -> B -> C of f need no grid barrier in between (f_rows_local); A*x still does
  const bool rows_local = f_rows_local(P);
  f_phase_pre(grid, P, W, x, sh, b, G, &ps);
  f_phase_A(P, W, x, sh, s_scr, b, G);
  if (hasA) gemv_n_phase(P.A, x, sh, b, G);
  if (hasA || !rows_local) grid.sync();
  if (rows_local) f_phase_B_local(P, W, x, s_scr, b, G); else f_phase_B(P, W, x, s_scr, b, G);
  if (hasA) for (int64_t i = tid; i < P.md; i += nt) W.Axb[axc][i] = zsum(P.A, i);
  if (!rows_local) grid.sync();
  f_phase_C(P, W, sh, b, G);
  if (hasA) gemv_t_phase(P.A, y, sh, b, G);
  grid.sync();
  {
    double tot[2];
    grid_totals<2>(W.red, G, SLOT_F0, tot, s_scr);
    grad_slice(P, W, j0, j1, W.gb[gc], tot[1], G);
    if (shardedF) p2p_allreduce<kThreads>(P.p2p, ps, grid, W.gb[gc], W.gb[gc], P.n);
    if (hasA) gsum_slice(P.A, j0, j1, W.Aty[atc], G);
    if (shardedA) p2p_allreduce<kThreads>(P.p2p, ps, grid, W.Aty[atc], W.Aty[atc], P.n);     // A'y over all row blocks
    primal_step(x, W.gb[gc], hasA ? W.Aty[atc] : nullptr, gamma, W.xb[1], gval_slot(1));     // :330-332
  }
  n_eval = 1; n_grad = 1; n_proxg = 1; n_mul = 1; n_amul = 1;
  grid.sync();
  double* x_prev = W.xb[0];
  x = W.xb[1]; xc = 1;
  double* grad_prev = W.gb[0];
  double norm_res = INFINITY;
  int64_t it_done = O.maxit;
  bool converged = false;

  for (int64_t it = 1; it <= O.maxit; ++it) {
    if (p2p_failed(P.p2p)) { flags |= ADAPROX_FLAG_COMM; it_done = it - 1; break; }   // a peer rank was lost (uniform: read after a grid barrier)
    // ---- P1 ---------------------------------------------------------------
    phase_stamp(W, it, 0);
    f_phase_pre(grid, P, W, x, sh, b, G, &ps);
    f_phase_A(P, W, x, sh, s_scr, b, G);                                        // :336
    if (hasA) gemv_n_phase(P.A, x, sh, b, G);                                   // :335
    if (hasA || !rows_local) grid.sync();
    // ---- P2 ---------------------------------------------------------------
    phase_stamp(W, it, 1);
    if (rows_local) f_phase_B_local(P, W, x, s_scr, b, G); else f_phase_B(P, W, x, s_scr, b, G);
    double* Ax_prev = W.Axb[axc];
    double* Ax = W.Axb[axc ^ 1];
    if (hasA) for (int64_t i = tid; i < P.md; i += nt) Ax[i] = zsum(P.A, i);
    n_eval++; n_mul++;
    if (!rows_local) grid.sync();
    // ---- P3 ---------------------------------------------------------------
    phase_stamp(W, it, 2);
    f_phase_C(P, W, sh, b, G);
    n_grad++;
    grid.sync();
    // ---- P4 ---------------------------------------------------------------
    phase_stamp(W, it, 3);
    double ftot[2];
    grid_totals<2>(W.red, G, SLOT_F0, ftot, s_scr);
    double* grad = W.gb[gc ^ 1];
    grad_slice(P, W, j0, j1, grad, ftot[1], G);
    if (shardedF) {
      p2p_allreduce<kThreads>(P.p2p, ps, grid, grad, grad, P.n);            // gather the gradient rows of all ranks
      p2p_allreduce_scalars<2>(P.p2p, ps, grid, ftot);                      // x'Qx and x'q over all rows
    }
    {
      double acc[4] = {0.0, 0.0, 0.0, 0.0};
      const double* aty = W.Aty[atc];
      for (int64_t j = j0 + threadIdx.x; j < j1; j += kThreads) {
        const double xj = x[j], gj = grad[j];
        const double pr = (W.v[j] - xj) / gamma + gj + (hasA ? aty[j] : 0.0);    // :338
        const double dg = gj - grad_prev[j], dx = xj - x_prev[j];
        acc[0] = fma(pr, pr, acc[0]);
        acc[1] = fma(dg, dg, acc[1]);
        acc[2] = fma(dg, dx, acc[2]);
        acc[3] = fma(dx, dx, acc[3]);
      }
      block_reduce_store<4>(acc, W.red, G, SLOT_PR, s_scr);
    }
    grid.sync();
    // ---- P5: stepsize, dual step ---------------------------------------------
    phase_stamp(W, it, 4);
    double t5[5];                                   // PR, GG, GX, DXX, GVAL
    {
      double t4[4], tg[1] = {0.0};
      grid_totals<4>(W.red, G, SLOT_PR, t4, s_scr);
      if (want_obj) grid_totals<1>(W.red, G, gval_slot(it), tg, s_scr);
      t5[0] = t4[0]; t5[1] = t4[1]; t5[2] = t4[2]; t5[3] = t4[3]; t5[4] = tg[0];
    }
    double xx0[1] = {0.0};
    if (P.f_kind == ADAPROX_F_CUBIC) grid_totals<1>(W.red, G, SLOT_XX0, xx0, s_scr);
    const double f_x = f_value(P, ftot[0], ftot[1], xx0[0]);
    const double g_x = (want_obj && !g_conj) ? prox_value_finish(P.g.kind, P.g.lambda, t5[4]) : NAN;   // a ConvexConjugate object is not callable
    const double gamma_old = gamma;                                              // :340
    double dr_sum = 0.0, h_sum = 0.0;
    double* ynew = y;
    double* w = W.w;

    if (!LINESEARCH) {
      rule_step(O, t5[1], t5[2], t5[3], gamma, sigma, s0, s1);                   // :341
      const double rho = gamma / gamma_old;                                      // :342
      if (hasA) {
        ynew = W.yb[yc ^ 1];
        double l2acc[1] = {0.0};
        for (int64_t i = tid; i < P.md; i += nt) {
          const double wi = y[i] + sigma * ((1.0 + rho) * Ax[i] - rho * Ax_prev[i]);   // :344
          w[i] = wi;
          if (h_l2) { const double z = prox_l2_arg(P.h, P.h.conjugate ? wi : wi / sigma, i); l2acc[0] = fma(z, z, l2acc[0]); }
        }
        double l2tot[1] = {0.0};
        if (h_l2) {
          block_reduce_store<1>(l2acc, W.red, G, SLOT_L2, s_scr);
          grid.sync();
          grid_totals<1>(W.red, G, SLOT_L2, l2tot, s_scr);
          if (shardedA) p2p_allreduce_scalars<1>(P.p2p, ps, grid, l2tot);                      // |w/sigma + shift|^2 over all rows
        }
        // each thread re-reads only the w[i] it wrote itself (same stride) -> no sync needed
        dual_rows(P, W, w, Ax, ynew, sigma, l2tot[0], want_obj, y, s_scr, b, G, false);
        n_proxh++;
        grid.sync();
        double t2[2];
        grid_totals<2>(W.red, G, SLOT_DR, t2, s_scr);
        if (!shardedA) {
          dr_sum = t2[0]; h_sum = t2[1];
        } else {
          // Row-sharded A: |dual_res|^2 and the h value are sums over ALL row blocks, and so is A'y+.  One exchange carries
          // the three of them (n + 2 doubles): A'y+ is formed here, BEFORE the convergence test that src/AdaProx.jl:348-358
          // runs first -- speculative work on the last iteration only (not counted: see n_amul below), one NVLink round trip
          // per iteration instead of two.
          gemv_t_phase(P.A, ynew, sh, b, G);                                          // :358 (speculative)
          grid.sync();
          gsum_slice(P.A, j0, j1, W.Aty[atc], G);
          if (b == 0 && threadIdx.x == 0) { W.Aty[atc][P.n] = t2[0]; W.Aty[atc][P.n + 1] = t2[1]; }
          grid.sync();
          p2p_allreduce<kThreads>(P.p2p, ps, grid, W.Aty[atc], W.Aty[atc], P.n + 2);
          dr_sum = ldcg(W.Aty[atc] + P.n); h_sum = ldcg(W.Aty[atc] + P.n + 1);
        }
      }
    } else {
      // ---- AdaPDM+ linesearch (:507-533) ----------------------------------------
      const double delta1 = 1.0 + O.delta;
      const double C = nan_to_zero(norm_sq_jl(t5[1]) / t5[2]);                   // :507
      const double L = nan_to_zero(t5[2] / norm_sq_jl(t5[3]));                   // :508
      const double Delta = gamma * L * (gamma * C - 1.0);                        // :509
      const double xi_bar = (O.t * O.t) * (gamma * gamma) * (eta * eta) * (delta1 * delta1);   // :510
      const double m4xim1 = 1.0 - 4.0 * xi_bar;                                  // :511
      eta = O.R * eta;                                                           // :513
      ynew = W.yb[yc ^ 1];
      double* aty_next = W.Aty[atc ^ 1];
      double gamma_next = gamma;
      double ls_dy = 0.0, ls_dr = 0.0, ls_h = 0.0;
      for (int trial = 0;; ++trial) {
        gamma_next = jl_min(jl_min(gamma * sqrt(1.0 + gamma / gamma_prev_ls), 1.0 / (2.0 * O.Theta * O.t * eta)),
                            gamma * sqrt(m4xim1 / (2.0 * delta1 * (Delta + sqrt(Delta * Delta + m4xim1 * sq(O.t * eta * gamma))))));   // :517-521
        const double rho = gamma_next / gamma;                                   // :522
        sigma = (O.t * O.t) * gamma_next;                                        // :523
        double l2acc[1] = {0.0};
        for (int64_t i = tid; i < P.md; i += nt) {
          const double wi = y[i] + sigma * ((1.0 + rho) * Ax[i] - rho * Ax_prev[i]);   // :524
          w[i] = wi;
          if (h_l2) { const double z = prox_l2_arg(P.h, P.h.conjugate ? wi : wi / sigma, i); l2acc[0] = fma(z, z, l2acc[0]); }
        }
        double l2tot[1] = {0.0};
        if (h_l2) {
          block_reduce_store<1>(l2acc, W.red, G, SLOT_L2, s_scr);
          grid.sync();
          grid_totals<1>(W.red, G, SLOT_L2, l2tot, s_scr);
          if (shardedA) p2p_allreduce_scalars<1>(P.p2p, ps, grid, l2tot);
        }
        dual_rows(P, W, w, Ax, ynew, sigma, l2tot[0], want_obj, y, s_scr, b, G, true);   // :525
        n_proxh++;
        grid.sync();
        gemv_t_phase(P.A, ynew, sh, b, G);                                           // :526
        n_amul++;
        grid.sync();
        gsum_slice(P.A, j0, j1, aty_next, G);
        if (shardedA) {
          // one exchange per trial: A'y_next and the three row sums of this trial (|y+ - y|^2, |dual_res|^2, h value)
          double td[1], tr[2];
          grid_totals<1>(W.red, G, SLOT_DY, td, s_scr);
          grid_totals<2>(W.red, G, SLOT_DR, tr, s_scr);
          if (b == 0 && threadIdx.x == 0) { aty_next[P.n] = td[0]; aty_next[P.n + 1] = tr[0]; aty_next[P.n + 2] = tr[1]; }
          grid.sync();
          p2p_allreduce<kThreads>(P.p2p, ps, grid, aty_next, aty_next, P.n + 3);
          ls_dy = ldcg(aty_next + P.n); ls_dr = ldcg(aty_next + P.n + 1); ls_h = ldcg(aty_next + P.n + 2);
        }
        {
          double acc[1] = {0.0};
          const double* aty = W.Aty[atc];
          for (int64_t j = j0 + threadIdx.x; j < j1; j += kThreads) {
            const double d = aty_next[j] - aty[j];
            acc[0] = fma(d, d, acc[0]);
          }
          block_reduce_store<1>(acc, W.red, G, SLOT_DATY, s_scr);
        }
        grid.sync();
        double tl[2];                               // DY, DATY
        grid_totals<2>(W.red, G, SLOT_DY, tl, s_scr);
        if (shardedA) tl[0] = ls_dy;                // |y_next - y|^2 is a sum over all row blocks; |A'y_next - A'y|^2 is already global
        const bool accept = eta >= sqrt(tl[1]) / sqrt(tl[0]);                    // :527
        if (accept || trial >= 200) {
          if (!accept) flags |= ADAPROX_FLAG_LS_CAP;
          gamma_prev_ls = gamma; gamma = gamma_next;                             // :528
          break;
        }
        eta *= O.r;                                                              // :532
        grid.sync();          // the retry rewrites reduction slots other CTAs may still be reading
      }
      double t2[2];
      grid_totals<2>(W.red, G, SLOT_DR, t2, s_scr);
      if (shardedA) { t2[0] = ls_dr; t2[1] = ls_h; }                             // from the accepted trial's exchange
      dr_sum = t2[0]; h_sum = t2[1];
      atc ^= 1;                                                                  // :529  At_y = At_y_next
    }

    // ---- P6: residual, record, convergence (:348-356) ----------------------------
    phase_stamp(W, it, 5);
    norm_res = sqrt(norm_sq_jl(t5[0]) + (hasA ? norm_sq_jl(dr_sum) : adapgm_dual_res_sq(gamma, gamma_old, sigma)));   // no dual vector: 0 or NaN (phases.cuh)
    if (!(gamma == gamma) || !(norm_res == norm_res) || isinf(gamma)) flags |= ADAPROX_FLAG_NONFINITE;
    if (b == 0 && threadIdx.x == 0 && W.rec != nullptr && it <= O.max_records) {
      adaprox_record rc;
      rc.it = it; rc.gamma = gamma; rc.sigma = sigma; rc.norm_res = norm_res;
      rc.f_x = f_x; rc.g_x = g_x;
      rc.h_Ax = (want_obj && hasA) ? (P.h.conjugate ? NAN : prox_value_finish(P.h.kind, P.h.lambda, h_sum)) : (want_obj ? 0.0 : NAN);
      rc.f_evals = n_eval; rc.grad_f_evals = n_grad; rc.prox_g_evals = n_proxg; rc.prox_h_evals = n_proxh;
      rc.A_evals = n_mul; rc.At_evals = n_amul;
      W.rec[it - 1] = rc;
    }
    if (it <= O.max_records) n_rec = it;
    if (hasA) { y = ynew; yc ^= 1; }
    if (norm_res <= O.tol) { converged = true; it_done = it; break; }            // :354-356

    if (hasA && !LINESEARCH) {
      n_amul++;
      if (!shardedA) {                                                               // sharded: already done with the scalars above
        gemv_t_phase(P.A, y, sh, b, G);                                              // :358
        grid.sync();
        gsum_slice(P.A, j0, j1, W.Aty[atc], G);
      }
    }
    // ---- P7 (:359-361) ---------------------------------------------------------------
    phase_stamp(W, it, 6);
    primal_step(x, grad, hasA ? W.Aty[atc] : nullptr, gamma, W.xb[(xc + 1) % 3], gval_slot(it + 1));   // :359-361; g(x+) is read in P5 of the next iteration
    n_proxg++;
    x_prev = x; xc = (xc + 1) % 3; x = W.xb[xc];                                 // :360
    grad_prev = grad; gc ^= 1;
    if (hasA) axc ^= 1;
    grid.sync();
    phase_stamp(W, it, 7);
  }

  // ---- epilogue: copy out -----------------------------------------------------
  for (int64_t j = tid; j < P.n; j += nt) W.xout[j] = x[j];
  if (hasA && W.yout) for (int64_t i = tid; i < P.md; i += nt) W.yout[i] = y[i];
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
