// solver_gridres.cuh -- AdaPGM / fixed-step PGM (src/AdaProx.jl:312-364 with A = 0, h = Zero) on a dense least-squares
// term whose matrix fits the shared memory of ALL SMs together (148 x ~225 KB = 33 MB), rows of at most 1024 columns.
//
// The reference's own lasso runs (lasso/runme.jl:191-195) are 100 x 300, 500 x 1000 and 4000 x 1000.  The first fits one
// cluster (solver_resident.cuh); the other two (4 MB, 32 MB) do not, and in the persistent grid kernel they spend their
// iteration on seven short phases (28 / 43 us per iteration: the matrix is L2-resident, the sweeps themselves are a few
// microseconds).  Here the grid keeps the matrix in shared memory for the whole solve, exactly as the cluster kernel does,
// exchanges through global memory (L2), and needs TWO grid barriers per iteration:
//   * CTA b owns rows [b R, (b+1) R) in its shared memory (loaded once) and a full copy of the iterate;
//   * pass 1 (one warp per row, x in registers): r_i = <A[i,:], x> - b_i;  pass 2 (thread = 2 columns, no reduction): this
//     CTA's partial gradient, written to its row of the matrix's partial buffer;
//   * grid barrier;  warp w of CTA b owns column b S + w: the partials of all row-owning CTAs in CTA order (each lane a fixed
//     subsequence, then the shuffle tree) -> one entry of the gradient in global memory;
//   * grid barrier;  EVERY CTA loads the whole gradient (8 KB) and repeats the rest of the iteration for all n columns: the
//     four stepsize sums of src/AdaProx.jl:338,260-261 (fixed order: the same bits in every CTA), the stepsize rule, the
//     convergence test, the prox step.  The new iterate never leaves the CTA, so there is no third barrier.
// MODE 1-4 run the comparison methods of the lasso experiment through the same passes and barriers (solver_pg.cuh holds their general
// grid form): fixed_nesterov (src/AdaProx.jl:91-142), agraal (:150-192), backtracking_proxgrad (:50-64), backtracking_nesterov (:66-84).
// A linesearch trial is a value-only pass 1 plus ONE grid barrier; the accepted trial's residual is reused for the pullback.
// No atomics, every sum has a fixed order: reruns are bit-identical.  Limits: ld <= 1024 (a row fits one warp's registers, a
// thread owns two columns), ceil(m / G) rows of A per CTA in shared memory, ceil(n / G) <= 16 gradient entries per CTA, G <= 256.
#pragma once
#include <cooperative_groups.h>

#include "phases.cuh"

namespace adaprox {

constexpr int kGThreads = 512;
constexpr int kGWarps = kGThreads / 32;
constexpr int kGMaxLd = 1024;
constexpr int kGLaneV = kGMaxLd / 2 / 32;                 // 16 double2 per lane per row (pass 1)
constexpr int kGMaxP = 8;                                 // partials per lane in a grid-wide sum: G <= 32 * kGMaxP
constexpr int kGSums = 5;                                 // |primal_res|^2, |dgrad|^2, <dgrad, dx>, |dx|^2, g(x)

struct GridResArgs {
  int rows_cap;            // rows per CTA: ceil(m / G)
  int slice;               // gradient entries reduced per CTA: ceil(n / G) <= kGWarps
  int row_ctas;            // CTAs that own at least one row: ceil(m / rows_cap)
  int x_in_smem;           // the CTA's copy of x lives in shared memory (else in xpriv: the last row did not leave room)
  double* gfull;           // [ld]     the gradient of the iteration
  double* fpart;           // [G]      per-CTA sums of r_i^2 (gradient evaluations)
  double* fpart2;          // [2][G]   the same for the value-only evaluations of MODE 1-4 (records, linesearch trials), alternating
  double* xpriv;           // [G][ld]  per-CTA copies of x when !x_in_smem
};

__host__ __device__ inline size_t gridres_smem_bytes(int64_t rows_cap, int64_t ld, bool x_in_smem) {
  return (size_t)8 * (size_t)(rows_cap * ld + 2 * rows_cap + kGWarps + kGWarps * 8 + 16 + (x_in_smem ? ld : 0));
}

template <int MODE>      // 0: AdaPGM / fixed-step PGM, 1: fixed_nesterov, 2: agraal, 3: backtracking_proxgrad, 4: backtracking_nesterov
__global__ void __launch_bounds__(kGThreads, 1) k_adapgm_gridres(DProblem P, DOpts O, DWork W, GridResArgs ga) {
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  extern __shared__ __align__(1024) unsigned char dyn_smem[];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int b = blockIdx.x, G = gridDim.x;
  const int64_t m = P.F.m, n = P.n, ld = P.F.ld;
  const int ldv = (int)(ld / 2);                           // double2 per row (<= 512)
  const int R = ga.rows_cap, S = ga.slice, GR = ga.row_ctas;
  const int64_t row0 = (int64_t)b * R;
  const int rows = (int)(row0 >= m ? 0 : (m - row0 < R ? m - row0 : R));
  double* As = reinterpret_cast<double*>(dyn_smem);       // [R][ld]
  double* r_loc = As + (size_t)R * ld;                     // [R]
  double* b_loc = r_loc + R;                               // [R]
  double* wpart = b_loc + R;                               // [kGWarps]      per-warp sums of r_i^2
  double* spart = wpart + kGWarps;                         // [kGWarps][8]   per-warp partials of the kGSums sums
  double* totals = spart + kGWarps * 8;                    // [8] the sums; [8] = sum of r_i^2 over the grid (CTA 0)
  double* xs = ga.x_in_smem ? totals + 16 : ga.xpriv + (int64_t)b * ld;    // [ld] this CTA's copy of the iterate
  double* gpart_all = P.F.gpart;                           // [G][npad] partial gradients
  const int64_t npad = P.F.npad;
  const bool want_obj = O.want_objective != 0;

  // thread t owns the columns c0 = 2 t and c0 + 1 in pass 2 and in the vector part of the iteration
  const int64_t c0 = 2 * (int64_t)t;
  const bool own0 = (t < ldv) && (c0 < n), own1 = (t < ldv) && (c0 + 1 < n);
  double2 x_t = make_double2(0.0, 0.0), xprev_t = x_t, gprev_t = x_t, v_t = x_t;

  // ---- load this CTA's rows, b and x0 ------------------------------------------------------------------------------
  {
    const double2* src = reinterpret_cast<const double2*>(P.F.a + row0 * ld);
    double2* dst = reinterpret_cast<double2*>(As);
    const int64_t cnt = (int64_t)rows * ldv;
    for (int64_t k = t; k < cnt; k += kGThreads) dst[k] = ld_stream(reinterpret_cast<const double*>(src + k));
    for (int i = t; i < R; i += kGThreads) { b_loc[i] = i < rows ? P.fvec[row0 + i] : 0.0; r_loc[i] = 0.0; }
    if (t < 16) totals[t] = 0.0;
    if (own0) x_t.x = W.xb[0][c0];
    if (own1) x_t.y = W.xb[0][c0 + 1];
    if (t < ldv) reinterpret_cast<double2*>(xs)[t] = x_t;             // columns n .. ld-1 hold zeros
  }
  __syncthreads();

  double gamma, sigma, s0, s1;
  rule_init(O, gamma, sigma, s0, s1);
  int64_t n_eval = 0, n_grad = 0, n_proxg = 0, n_rec = 0;
  unsigned flags = 0;
  double norm_res = INFINITY;
  int64_t it_done = O.maxit;
  bool converged = false;

  // pass 1 at xs (one warp per row, x in registers): r_loc[i] = <A[i,:], x> - b_i, this CTA's sum of r_i^2 -> fdst[b]
  auto pass1 = [&](double* fdst) {
    double2 xr[kGLaneV];
#pragma unroll
    for (int k = 0; k < kGLaneV; ++k) {
      const int idx = lane + 32 * k;
      xr[k] = idx < ldv ? reinterpret_cast<const double2*>(xs)[idx] : make_double2(0.0, 0.0);
    }
    double fw = 0.0;
    for (int i = warp; i < rows; i += kGWarps) {
      const double2* row = reinterpret_cast<const double2*>(As + (size_t)i * ld);
      double p0 = 0.0, p1 = 0.0;
#pragma unroll
      for (int k = 0; k < kGLaneV; ++k) {
        const int idx = lane + 32 * k;
        if (idx < ldv) { const double2 a = row[idx]; p0 = fma(a.x, xr[k].x, p0); p1 = fma(a.y, xr[k].y, p1); }
      }
      const double res = warp_sum(p0 + p1) - b_loc[i];                          // lasso/runme.jl:22
      if (lane == 0) r_loc[i] = res;
      fw = fma(res, res, fw);
    }
    if (lane == 0) wpart[warp] = fw;
    __syncthreads();
    if (t == 0) {
      double fs = 0.0;
#pragma unroll
      for (int w = 0; w < kGWarps; ++w) fs += wpart[w];
      fdst[b] = fs;
    }
  };
  // pass 2 from r_loc (thread t owns double2 column t, no reduction): this CTA's partial gradient -> gpart_all[b][:]
  auto pass2 = [&]() {
    if (t < ldv && b < GR) {
      double2 acc = make_double2(0.0, 0.0);
      for (int i = 0; i < rows; ++i) {
        const double ri = r_loc[i];
        const double2 a = reinterpret_cast<const double2*>(As + (size_t)i * ld)[t];
        acc.x = fma(a.x, ri, acc.x); acc.y = fma(a.y, ri, acc.y);              // :23
      }
      *reinterpret_cast<double2*>(gpart_all + (int64_t)b * npad + c0) = acc;
    }
  };
  // value + pullback at xs; value partial -> fpart[b]
  auto local_gradient = [&]() { pass1(ga.fpart); pass2(); };
  // value-only evaluations (records of MODE 1 / 2, linesearch trials of MODE 3 / 4) alternate between two partial buffers: a CTA
  // may start the next one while another still reads this one's partials (one grid barrier apart, not two)
  unsigned vcount = 0;
  // last warp (of CTA 0, or of every CTA): the G value partials in CTA order -> totals[8]
  auto value_total = [&](const double* part, bool all) {
    if ((all || b == 0) && warp == kGWarps - 1) {
      double v[kGMaxP];
#pragma unroll
      for (int q = 0; q < kGMaxP; ++q) { const int p = lane + 32 * q; v[q] = p < G ? ldcg(part + p) : 0.0; }
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < kGMaxP; ++q) s += v[q];
      s = warp_sum(s);
      if (lane == 0) totals[8] = s;
    }
  };
  // after the first barrier: warp w reduces gradient entry b S + w (lane l the CTAs l, l + 32, ... in order, then the shuffle
  // tree); the last warp of CTA 0 sums the value partials for the record the same way
  auto reduce_slice = [&]() {
    const int64_t j = (int64_t)b * S + warp;
    if (warp < S && j < n) {
      double v[kGMaxP];
#pragma unroll
      for (int q = 0; q < kGMaxP; ++q) { const int p = lane + 32 * q; v[q] = p < GR ? ldcg(gpart_all + (int64_t)p * npad + j) : 0.0; }
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < kGMaxP; ++q) s += v[q];
      s = warp_sum(s);
      if (lane == 0) ga.gfull[j] = s;
    }
    if (MODE == 0) value_total(ga.fpart, false);
    if (MODE == 4) value_total(ga.fpart, true);
  };
  // after the second barrier: this thread's two gradient entries
  auto load_gradient = [&]() -> double2 {
    double2 g = make_double2(0.0, 0.0);
    if (t < ldv) { g = ldcg2(ga.gfull + c0); if (!own0) g.x = 0.0; if (!own1) g.y = 0.0; }
    return g;
  };
  // kGSums per-thread terms -> totals[0 .. kGSums): warp tree, then the warps in order; identical in every CTA
  auto block_sums = [&](const double (&a)[kGSums]) {
#pragma unroll
    for (int k = 0; k < kGSums; ++k) {
      const double s = warp_sum(a[k]);
      if (lane == 0) spart[warp * 8 + k] = s;
    }
    __syncthreads();
    if (t < kGSums) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < kGWarps; ++w) s += spart[w * 8 + t];
      totals[t] = s;
    }
    __syncthreads();
  };
  auto prox_step = [&](const double2 g) {
    if (t < ldv) {
      v_t.x = x_t.x - gamma * g.x; v_t.y = x_t.y - gamma * g.y;                 // :330 / :359
      double2 xn = make_double2(0.0, 0.0);
      if (own0) xn.x = prox_elem(P.g, v_t.x, gamma, c0, 0.0);                   // :332 / :361
      if (own1) xn.y = prox_elem(P.g, v_t.y, gamma, c0 + 1, 0.0);
      xprev_t = x_t; gprev_t = g; x_t = xn;
      reinterpret_cast<double2*>(xs)[t] = xn;
    }
    __syncthreads();                                       // xs complete (shared memory, or this CTA's own global copy)
  };

  if constexpr (MODE == 0) {
    // ---- prologue (:327-332) ---------------------------------------------------------------------------------------------
    {
      local_gradient();
      grid.sync();                                          // all partial gradients are in place
      reduce_slice();
      grid.sync();                                          // the gradient is complete
      n_eval = 1; n_grad = 1;
      prox_step(load_gradient());
      n_proxg = 1;
    }

    for (int64_t it = 1; it <= O.maxit; ++it) {
      phase_stamp(W, it, 0);                                // ADAPROX_PHASE_TIMING=1: CTA 0's clock at the phase boundaries
      local_gradient();                                     // :336 value + pullback
      n_eval++; n_grad++;
      phase_stamp(W, it, 1);
      grid.sync();                                          // barrier 1
      phase_stamp(W, it, 2);
      reduce_slice();
      phase_stamp(W, it, 3);
      grid.sync();                                          // barrier 2
      phase_stamp(W, it, 4);
      const double2 g = load_gradient();
      {
        double a[kGSums] = {0.0, 0.0, 0.0, 0.0, 0.0};
        if (own0) {
          const double pr = (v_t.x - x_t.x) / gamma + g.x;                        // :338 (old gamma)
          const double dg = g.x - gprev_t.x, dx = x_t.x - xprev_t.x;
          a[0] = pr * pr; a[1] = dg * dg; a[2] = dg * dx; a[3] = dx * dx;
          if (want_obj) a[4] = prox_value_elem(P.g, x_t.x, c0);
        }
        if (own1) {
          const double pr = (v_t.y - x_t.y) / gamma + g.y;
          const double dg = g.y - gprev_t.y, dx = x_t.y - xprev_t.y;
          a[0] = fma(pr, pr, a[0]); a[1] = fma(dg, dg, a[1]); a[2] = fma(dg, dx, a[2]); a[3] = fma(dx, dx, a[3]);
          if (want_obj) a[4] += prox_value_elem(P.g, x_t.y, c0 + 1);
        }
        // Warp 0 leaves cooperative_groups' grid.sync() diverged (its thread 0 polled the arrival counter), and ptxas cannot prove
        // convergence across the predicated blocks above: without a convergence point here the 50 SHFL of the five shuffle trees are
        // emitted a second time behind BRA.DIV / WARPSYNC guards (SASS: 224 SHFL, 121 WARPSYNC instead of 174 / 71) and the diverged
        // warp takes that path -- 6.2 us for this phase instead of 1.2 us (500 x 1000: 15.7 vs 8.2 us per iteration, measured;
        // __syncwarp and a CTA barrier do equally well; -DADAPROX_EXP_GR_NO_PRESYNC removes it for the A/B).
  #ifndef ADAPROX_EXP_GR_NO_PRESYNC
        __syncwarp();
  #endif
        block_sums(a);
      }
      phase_stamp(W, it, 5);
      const double gamma_prev = gamma;
      rule_step(O, totals[1], totals[2], totals[3], gamma, sigma, s0, s1);        // :341
      norm_res = sqrt(norm_sq_jl(totals[0]) + adapgm_dual_res_sq(gamma, gamma_prev, sigma));   // :348 (dual part: 0, or NaN -- phases.cuh)
      if (!(gamma == gamma) || !(norm_res == norm_res) || isinf(gamma)) flags |= ADAPROX_FLAG_NONFINITE;
      if (b == 0 && t == 0 && W.rec != nullptr && it <= O.max_records) {
        adaprox_record rc;
        rc.it = it; rc.gamma = gamma; rc.sigma = sigma; rc.norm_res = norm_res;
        rc.f_x = 0.5 * norm_sq_jl(totals[8]);
        rc.g_x = want_obj ? prox_value_finish(P.g.kind, P.g.lambda, totals[4]) : NAN;
        rc.h_Ax = want_obj ? 0.0 : NAN;
        rc.f_evals = n_eval; rc.grad_f_evals = n_grad; rc.prox_g_evals = n_proxg; rc.prox_h_evals = 0;
        rc.A_evals = 0; rc.At_evals = 0;
        W.rec[it - 1] = rc;
      }
      if (it <= O.max_records) n_rec = it;
      if (norm_res <= O.tol) { converged = true; it_done = it; break; }           // :354-356 (uniform over the grid: same bits everywhere)
      phase_stamp(W, it, 6);
      prox_step(g);
      n_proxg++;
      phase_stamp(W, it, 7);
    }
  }

  // ---- the two comparison methods: shared pieces --------------------------------------------------------------------------
  // full gradient at the point the owning threads hold in `at`: xs <- at, passes, two grid barriers, this thread's two entries
  auto gradient_at = [&](const double2 at) -> double2 {
    if (t < ldv) reinterpret_cast<double2*>(xs)[t] = at;
    __syncthreads();
    local_gradient();
    grid.sync();
    reduce_slice();
    grid.sync();
    return load_gradient();
  };
  // f at the point in `at`, NOT counted (the objective of a record: `without_counting`, src/AdaProx.jl:134-136 / :183-185);
  // one more grid barrier.  `all`: every CTA obtains the value (linesearch), else only CTA 0, which writes the records.
  auto value_at = [&](const double2 at, bool all) -> double {
    double* part = ga.fpart2 + (size_t)(vcount++ & 1u) * G;
    if (t < ldv) reinterpret_cast<double2*>(xs)[t] = at;
    __syncthreads();
    pass1(part);
    grid.sync();
    value_total(part, all);
    __syncthreads();
    return f_value(P, totals[8], 0.0, 0.0);
  };
  auto record = [&](int64_t it, double objective_f, double objective_g) {
    if (b == 0 && t == 0 && W.rec != nullptr && it <= O.max_records) {
      adaprox_record rc;
      rc.it = it; rc.gamma = gamma; rc.sigma = NAN; rc.norm_res = norm_res;
      rc.f_x = objective_f; rc.g_x = objective_g; rc.h_Ax = 0.0;
      rc.f_evals = n_eval; rc.grad_f_evals = n_grad; rc.prox_g_evals = n_proxg; rc.prox_h_evals = 0;
      rc.A_evals = 0; rc.At_evals = 0;
      W.rec[it - 1] = rc;
    }
    if (it <= O.max_records) n_rec = it;
  };
  // prox of (p - gamma g) for this thread's two columns; adds |prox - ref|^2 to a[0] and g(prox) to a[1]
  auto prox_pair = [&](const double2 p, const double2 g, const double2 ref, double (&a)[kGSums]) -> double2 {
    double2 xn = make_double2(0.0, 0.0);
    if (own0) {
      xn.x = prox_elem(P.g, p.x - gamma * g.x, gamma, c0, 0.0);
      const double d = xn.x - ref.x;
      a[0] = fma(d, d, a[0]);
      a[1] += prox_value_elem(P.g, xn.x, c0);
    }
    if (own1) {
      xn.y = prox_elem(P.g, p.y - gamma * g.y, gamma, c0 + 1, 0.0);
      const double d = xn.y - ref.y;
      a[0] = fma(d, d, a[0]);
      a[1] += prox_value_elem(P.g, xn.y, c0 + 1);
    }
    return xn;
  };

  if constexpr (MODE == 1) {                               // fixed_nesterov, src/AdaProx.jl:91-142 (grid form: solver_pg.cuh)
    sigma = NAN;
    gamma = O.gamma;
    const double mu = O.muf + O.mug;                                             // :108-117
    const double q = gamma * mu / (1.0 + gamma * O.mug);
    double theta = O.theta >= 0.0 ? O.theta : (q > 0.0 ? 1.0 / sqrt(q) : 0.0);
    xprev_t = x_t;                                                               // :119
    for (int64_t it = 1; it <= O.maxit; ++it) {
      const double theta_prev = theta;
      double beta;
      if (mu == 0.0) {                                                           // :122-128
        theta = (1.0 + sqrt(1.0 + 4.0 * theta_prev * theta_prev)) / 2.0;
        beta = (theta_prev - 1.0) / theta;
      } else {
        const double a_ = 1.0 - q * theta_prev * theta_prev;
        theta = (a_ + sqrt(a_ * a_ + 4.0 * theta_prev * theta_prev)) / 2.0;
        beta = (theta_prev - 1.0) * (1.0 + gamma * O.mug - theta * gamma * mu) / theta / (1.0 - gamma * O.muf);
      }
      double2 z = make_double2(0.0, 0.0);
      if (own0) z.x = x_t.x + beta * (x_t.x - xprev_t.x);                        // :129
      if (own1) z.y = x_t.y + beta * (x_t.y - xprev_t.y);
      const double2 g = gradient_at(z);                                          // :130
      n_eval++; n_grad++;
      double a[kGSums] = {0.0, 0.0, 0.0, 0.0, 0.0};
      const double2 xn = prox_pair(z, g, z, a);                                  // :131-133
      __syncwarp();
      block_sums(a);
      n_proxg++;
      xprev_t = x_t; x_t = xn;
      norm_res = sqrt(totals[0]) / gamma;
      const double g_x = prox_value_finish(P.g.kind, P.g.lambda, totals[1]);
      const double fx = want_obj ? value_at(x_t, false) : NAN;                   // :134-136
      record(it, fx, g_x);
      if (norm_res <= O.tol) { converged = true; it_done = it; break; }
    }
    if (!(gamma == gamma) || !(norm_res == norm_res)) flags |= ADAPROX_FLAG_NONFINITE;
  }

  if constexpr (MODE == 2) {                               // agraal, src/AdaProx.jl:150-192; W.aux[0] holds the second start point x0
    sigma = NAN;
    gamma = O.gamma;
    double2 xbar_t = x_t, g_t, gp_t;
    xprev_t = make_double2(own0 ? W.aux[0][c0] : 0.0, own1 ? W.aux[0][c0 + 1] : 0.0);   // :165
    g_t = gradient_at(x_t);                                                      // :166
    gp_t = gradient_at(xprev_t);                                                 // :167
    n_eval = 2; n_grad = 2;
    const double phi = O.phi;
    const double rho = 1.0 / phi + 1.0 / (phi * phi);                            // :172
    double theta = 1.0;
    for (int64_t it = 1; it <= O.maxit; ++it) {
      double a[kGSums] = {0.0, 0.0, 0.0, 0.0, 0.0};
      if (own0) { const double dx = x_t.x - xprev_t.x, dg = g_t.x - gp_t.x; a[0] = dx * dx; a[1] = dg * dg; }
      if (own1) { const double dx = x_t.y - xprev_t.y, dg = g_t.y - gp_t.y; a[0] = fma(dx, dx, a[0]); a[1] = fma(dg, dg, a[1]); }
      __syncwarp();
      block_sums(a);
      const double sxx = totals[0], sgg = totals[1];
      if (it == 1 && !(O.gamma > 0.0)) gamma = sqrt(sxx) / sqrt(sgg);            // :168-170
      const double C = norm_sq_jl(sxx) / norm_sq_jl(sgg);                        // :175
      const double gamma_prev = gamma;
      gamma = jl_min(jl_min(rho * gamma_prev, phi * theta * C / (4.0 * gamma_prev)), O.gamma_max);   // :177
      theta = phi * gamma / gamma_prev;                                          // :178
      double2 xb = make_double2(0.0, 0.0);
      if (own0) xb.x = ((phi - 1.0) * x_t.x + xbar_t.x) / phi;                   // :179
      if (own1) xb.y = ((phi - 1.0) * x_t.y + xbar_t.y) / phi;
      xbar_t = xb;
      double a2[kGSums] = {0.0, 0.0, 0.0, 0.0, 0.0};
      const double2 xn = prox_pair(xb, g_t, x_t, a2);                            // :181
      __syncwarp();
      block_sums(a2);
      n_proxg++;
      xprev_t = x_t; x_t = xn; gp_t = g_t;                                       // :180
      norm_res = sqrt(totals[0]) / gamma;                                        // :182
      const double g_x = prox_value_finish(P.g.kind, P.g.lambda, totals[1]);
      const double fx = want_obj ? value_at(x_t, false) : NAN;                   // :183-185
      record(it, fx, g_x);
      if (norm_res <= O.tol) { converged = true; it_done = it; break; }
      g_t = gradient_at(x_t);                                                    // :189
      n_eval++; n_grad++;
    }
    if (!(gamma == gamma) || !(norm_res == norm_res)) flags |= ADAPROX_FLAG_NONFINITE;
  }

  double2 res_t = x_t;                                     // MODE 3 / 4: the last accepted point
  if constexpr (MODE == 3 || MODE == 4) {                  // backtracking_proxgrad (:50-64) / backtracking_nesterov (:66-84); grid form: solver_pg.cuh
    constexpr bool nesterov = (MODE == 4);
    sigma = NAN;
    gamma = O.gamma;
    double2 zprev_t = x_t;                                                       // :67
    double theta = 1.0;                                                          // :68
    double2 g_t = gradient_at(x_t);                                              // :52 / :69
    // f(x) from the same evaluation: every CTA needs it for the sufficient-decrease test
    if (!nesterov) value_total(ga.fpart, true);                                  // (MODE 4: reduce_slice already did)
    __syncthreads();
    double f_x = f_value(P, totals[8], 0.0, 0.0);
    n_eval = 1; n_grad = 1;
    for (int64_t it = 1; it <= O.maxit; ++it) {
      gamma = nesterov ? gamma : O.xi * gamma;                                   // :54 / :72
      double f_z = 0.0, g_z = 0.0, dzz = 0.0;
      double2 z = make_double2(0.0, 0.0);
      for (;;) {                                                                 // backtrack_stepsize (:34-48)
        double a[kGSums] = {0.0, 0.0, 0.0, 0.0, 0.0};
        z = make_double2(0.0, 0.0);
        if (own0) {
          z.x = prox_elem(P.g, x_t.x - gamma * g_t.x, gamma, c0, 0.0);           // :35 / :43
          const double d = z.x - x_t.x;
          a[0] = g_t.x * d; a[1] = d * d; a[2] = prox_value_elem(P.g, z.x, c0);
        }
        if (own1) {
          z.y = prox_elem(P.g, x_t.y - gamma * g_t.y, gamma, c0 + 1, 0.0);
          const double d = z.y - x_t.y;
          a[0] = fma(g_t.y, d, a[0]); a[1] = fma(d, d, a[1]); a[2] += prox_value_elem(P.g, z.y, c0 + 1);
        }
        __syncwarp();
        block_sums(a);
        n_proxg++;
        const double gd = totals[0];
        dzz = totals[1];
        g_z = prox_value_finish(P.g.kind, P.g.lambda, totals[2]);
        f_z = value_at(z, true);                                                 // :37 / :45
        n_eval++;
        const double ub_z = f_x + gd + 1.0 / (2.0 * gamma) * norm_sq_jl(dzz);    // :26
        if (!(f_z > ub_z)) break;                                                // :38
        gamma *= O.shrink;                                                       // :39
        if (gamma < 1e-12) flags |= ADAPROX_FLAG_STEP_TOO_SMALL;                 // :40-42 (the reference keeps looping)
        if (gamma < 1e-300) break;
      }
      norm_res = sqrt(dzz) / gamma;                                              // :55 / :73
      record(it, f_z, g_z);
      res_t = z;
      if (norm_res <= O.tol) { converged = true; it_done = it; break; }          // :57 / :75
      if (!nesterov) {
        pass2();                                                                 // :60-61  grad_x = pb(): r_loc still holds the residual of z
        grid.sync();
        reduce_slice();
        grid.sync();
        g_t = load_gradient();
        n_grad++;
        x_t = z; f_x = f_z;
      } else {
        const double theta_prev = theta;                                         // :78-80
        theta = (1.0 + sqrt(1.0 + 4.0 * theta_prev * theta_prev)) / 2.0;
        const double beta = (theta_prev - 1.0) / theta;
        if (own0) x_t.x = z.x + beta * (z.x - zprev_t.x);
        if (own1) x_t.y = z.y + beta * (z.y - zprev_t.y);
        zprev_t = z;                                                             // :71
        g_t = gradient_at(x_t);                                                  // :81 (reduce_slice leaves f(x) in totals[8] of every CTA)
        __syncthreads();
        f_x = f_value(P, totals[8], 0.0, 0.0);
        n_eval++; n_grad++;
      }
    }
    if (!(gamma == gamma) || !(norm_res == norm_res)) flags |= ADAPROX_FLAG_NONFINITE;
    x_t = res_t;
  }

  if (b == 0) {                                            // converged: the iterate whose gradient was just evaluated; maxit: the last prox
    if (own0) W.xout[c0] = x_t.x;
    if (own1) W.xout[c0 + 1] = x_t.y;
  }
  if (b == 0 && t == 0) {
    DResult r;
    r.iters = it_done;
    r.flags = flags | (converged ? ADAPROX_FLAG_CONVERGED : 0u);
    r.xbuf = 0;
    r.f_evals = n_eval; r.grad_f_evals = n_grad; r.prox_g_evals = n_proxg; r.prox_h_evals = 0;
    r.A_evals = 0; r.At_evals = 0; r.n_records = n_rec;
    r.final_gamma = gamma; r.final_sigma = sigma; r.final_norm_res = norm_res;
    *W.res = r;
  }
}

}  // namespace adaprox
