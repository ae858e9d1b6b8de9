// solver_resident.cuh -- AdaPGM / fixed-step PGM (src/AdaProx.jl:312-364 with A = 0, h = Zero) on a SMALL dense
// least-squares term: the matrix lives in the distributed shared memory of ONE 16-CTA cluster for the whole solve.
//
// Every lasso instance the reference's own experiments run on a laptop (lasso/runme.jl:192-207: 100 x 300, 400 x 1000)
// is a few megabytes: the persistent grid kernel (solver_pd.cuh) spends such an iteration on grid barriers (1.2 us each,
// profiles/r02_latency_probe.jsonl) and on L2 round trips of per-CTA partials -- 25 us per iteration at 400 x 1000.  Here
//   * CTA c of the cluster owns rows [c R, (c+1) R) of A in its shared memory (loaded once: 16 x ~200 KB = 3.2 MB);
//   * pass 1 (one warp per row, x in registers): r_i = <A[i,:], x> - b_i;
//   * pass 2 (thread = 4 columns, no reduction at all): this CTA's partial gradient  gp[j] = sum_i A[i,j] r_i;
//   * hardware cluster barrier (0.23 us);  CTA c sums columns [c S, (c+1) S) of the 16 partial gradients over DSMEM in rank
//     order, forms the four stepsize reductions of src/AdaProx.jl:338,260-261 for its slice;
//   * cluster barrier;  every CTA sums the 16 x 6 scalars in rank order (the same bits everywhere), applies the stepsize rule,
//     the convergence test and the prox step on its slice, and writes its new x entries into every peer's copy of x;
//   * cluster barrier.
// Three cluster barriers per iteration, no global memory on the path, no atomics; every sum has a fixed order, so reruns are
// bit-identical.  Limits: n <= 1024 (a row fits one warp's registers), m * ld * 8 <= 16 x (227 KB - vectors).
#pragma once
#include "phases.cuh"
#include "solver_fused.cuh"      // cluster helpers

namespace adaprox {

constexpr int kRThreads = 256;
constexpr int kRWarps = kRThreads / 32;
constexpr int kRCluster = 16;
constexpr int kRMaxLd = 1024;
constexpr int kRLaneV = kRMaxLd / 2 / 32;                 // 16 double2 per lane per row (pass 1)
constexpr int kRThrV = kRMaxLd / 2 / kRThreads;           // 2 double2 per thread per row (pass 2)
constexpr int kRScal = 8;                                 // scalars exchanged per CTA per iteration

struct ResidentArgs {
  int rows_cap;            // rows per CTA (ceil(m / 16))
  int slice;               // columns per CTA in the reduce / prox phase (ceil(n / 16))
};

__host__ __device__ inline size_t resident_smem_bytes(int64_t rows_cap, int64_t ld) {
  return (size_t)8 * (size_t)(rows_cap * ld + 2 * ld + 2 * rows_cap + 2 * kRScal + 4 * kRWarps);
}

__device__ __forceinline__ double ld_dsmem(uint32_t local_addr, uint32_t peer) {
  uint32_t ra; double v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(peer));
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra) : "memory");
  return v;
}
__device__ __forceinline__ void st_dsmem(uint32_t local_addr, uint32_t peer, double v) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(peer));
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(ra), "d"(v) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() { cluster_arrive(); cluster_wait(); }

__global__ void __launch_bounds__(kRThreads, 1) k_adapgm_resident(DProblem P, DOpts O, DWork W, ResidentArgs ra) {
  extern __shared__ __align__(1024) unsigned char dyn_smem[];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const uint32_t rank = cluster_ctarank();
  const int64_t m = P.F.m, n = P.n, ld = P.F.ld;
  const int ldv = (int)(ld / 2);                           // double2 per row
  const int R = ra.rows_cap, S = ra.slice;
  const int64_t row0 = (int64_t)rank * R;
  const int rows = (int)(row0 >= m ? 0 : (m - row0 < R ? m - row0 : R));
  double* As = reinterpret_cast<double*>(dyn_smem);       // [R][ld]
  double* xs = As + (size_t)R * ld;                        // [ld]  current iterate, full copy
  double* gp = xs + ld;                                    // [ld]  this CTA's partial gradient
  double* r_loc = gp + ld;                                 // [R]
  double* b_loc = r_loc + R;                               // [R]
  double* exch = b_loc + R;                                // [kRScal] this CTA's scalars of the iteration
  double* totals = exch + kRScal;                          // [kRScal] sums over the cluster
  double* wpart = totals + kRScal;                         // [4 * kRWarps] per-warp partials
  const uint32_t xs_a = smem_u32(xs), gp_a = smem_u32(gp), exch_a = smem_u32(exch);
  const bool want_obj = O.want_objective != 0;

  // ---- load this CTA's rows, b, x0 ---------------------------------------------------------------------------------
  {
    const double2* src = reinterpret_cast<const double2*>(P.F.a + row0 * ld);
    double2* dst = reinterpret_cast<double2*>(As);
    const int64_t cnt = (int64_t)rows * ldv;
    for (int64_t k = t; k < cnt; k += kRThreads) dst[k] = src[k];
    for (int i = t; i < R; i += kRThreads) { b_loc[i] = i < rows ? P.fvec[row0 + i] : 0.0; r_loc[i] = 0.0; }
    for (int64_t j = t; j < ld; j += kRThreads) { xs[j] = j < n ? W.xb[0][j] : 0.0; gp[j] = 0.0; }
    if (t < kRScal) { exch[t] = 0.0; totals[t] = 0.0; }
  }
  __syncthreads();
  cluster_sync_all();

  // slice state of thread t (column j of this CTA's slice): x_j, x_prev_j, grad_prev_j, v_j
  const int64_t j = (int64_t)rank * S + t;
  const bool own = (t < S) && (j < n);
  double x_j = own ? xs[j] : 0.0, xprev_j = 0.0, gprev_j = 0.0, v_j = 0.0;

  double gamma, sigma, s0, s1;
  rule_init(O, gamma, sigma, s0, s1);
  int64_t n_eval = 0, n_grad = 0, n_proxg = 0, n_rec = 0;
  unsigned flags = 0;
  double norm_res = INFINITY;
  int64_t it_done = O.maxit;
  bool converged = false;
  double gval_part = 0.0;                                  // g(x) partial of this CTA's slice for the NEXT record

  // gradient evaluation: leaves gp (partial gradient of this CTA's rows) and returns this CTA's sum of r_i^2
  auto local_gradient = [&]() -> double {
    // pass 1: one warp per row, x in registers
    double2 xr[kRLaneV];
#pragma unroll
    for (int k = 0; k < kRLaneV; ++k) {
      const int idx = lane + 32 * k;
      xr[k] = idx < ldv ? reinterpret_cast<const double2*>(xs)[idx] : make_double2(0.0, 0.0);
    }
    double fw = 0.0;
    for (int i = warp; i < rows; i += kRWarps) {
      const double2* row = reinterpret_cast<const double2*>(As + (size_t)i * ld);
      double p0 = 0.0, p1 = 0.0;
#pragma unroll
      for (int k = 0; k < kRLaneV; ++k) {
        const int idx = lane + 32 * k;
        if (idx < ldv) { const double2 a = row[idx]; p0 = fma(a.x, xr[k].x, p0); p1 = fma(a.y, xr[k].y, p1); }
      }
      const double res = warp_sum(p0 + p1) - b_loc[i];                          // lasso/runme.jl:22
      if (lane == 0) r_loc[i] = res;
      fw = fma(res, res, fw);
    }
    if (lane == 0) wpart[warp] = fw;
    __syncthreads();
    // pass 2: thread t owns the double2 columns t and t + 256: no reduction
    double2 acc[kRThrV];
#pragma unroll
    for (int k = 0; k < kRThrV; ++k) acc[k] = make_double2(0.0, 0.0);
    for (int i = 0; i < rows; ++i) {
      const double ri = r_loc[i];
      const double2* row = reinterpret_cast<const double2*>(As + (size_t)i * ld);
#pragma unroll
      for (int k = 0; k < kRThrV; ++k) {
        const int idx = t + kRThreads * k;
        if (idx < ldv) { const double2 a = row[idx]; acc[k].x = fma(a.x, ri, acc[k].x); acc[k].y = fma(a.y, ri, acc[k].y); }   // :23
      }
    }
#pragma unroll
    for (int k = 0; k < kRThrV; ++k) {
      const int idx = t + kRThreads * k;
      if (idx < ldv) reinterpret_cast<double2*>(gp)[idx] = acc[k];
    }
    double fs = 0.0;
#pragma unroll
    for (int w = 0; w < kRWarps; ++w) fs += wpart[w];
    return fs;                                            // same value in every thread of the CTA
  };
  // gradient entry j of this CTA's slice: the 16 partials in rank order (DSMEM)
  auto reduce_slice = [&]() -> double {
    double v[kRCluster];
#pragma unroll
    for (int p = 0; p < kRCluster; ++p) v[p] = own ? ld_dsmem(gp_a + (uint32_t)j * 8, (uint32_t)p) : 0.0;
    double s = 0.0;
#pragma unroll
    for (int p = 0; p < kRCluster; ++p) s += v[p];
    return s;
  };
  // K slice-partials (threads < S) -> exch[k0 .. k0 + K)
  auto slice_sums = [&](double (&a)[4], int K, int k0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < K) { const double s = warp_sum(a[k]); if (lane == 0) wpart[warp * 4 + k] = s; }
    }
    __syncthreads();
    if (t < K) {
      double s = 0.0;
      const int nw = (S + 31) / 32;                       // warps that hold slice columns
      for (int w = 0; w < nw; ++w) s += wpart[w * 4 + t];
      exch[k0 + t] = s;
    }
    __syncthreads();
  };
  // sum the cluster's scalars in rank order: totals[k], identical bits in every CTA
  auto cluster_totals = [&]() {
    if (t < kRScal) {
      double s = 0.0;
#pragma unroll
      for (int p = 0; p < kRCluster; ++p) s += ld_dsmem(exch_a + t * 8, (uint32_t)p);
      totals[t] = s;
    }
    __syncthreads();
  };
  // prox step on the slice, broadcast of the new entries into every peer's x
  auto prox_and_broadcast = [&](double gj) {
    if (own) {
      v_j = x_j - gamma * gj;                                                   // :330 / :359
      const double xn = prox_elem(P.g, v_j, gamma, j, 0.0);                     // :332 / :361
      xprev_j = x_j; gprev_j = gj; x_j = xn;
      gval_part = want_obj ? prox_value_elem(P.g, xn, j) : 0.0;
#pragma unroll
      for (int p = 0; p < kRCluster; ++p) st_dsmem(xs_a + (uint32_t)j * 8, (uint32_t)p, xn);
    } else {
      gval_part = 0.0;
    }
  };

  // ---- prologue (:327-332) ---------------------------------------------------------------------------------------------
  {
    (void)local_gradient();
    cluster_sync_all();                                   // all partial gradients are in place
    const double gj = reduce_slice();
    n_eval = 1; n_grad = 1;
    cluster_sync_all();                                   // everybody has read gp and xs: both may be overwritten
    prox_and_broadcast(gj);
    n_proxg = 1;
    cluster_sync_all();
  }

  for (int64_t it = 1; it <= O.maxit; ++it) {
    const double fs = local_gradient();                   // :336 value + pullback
    n_eval++; n_grad++;
    cluster_sync_all();                                   // barrier 1
    const double gj = reduce_slice();
    {
      double a[4] = {0.0, 0.0, 0.0, 0.0};
      if (own) {
        const double pr = (v_j - x_j) / gamma + gj;                             // :338 (old gamma)
        const double dg = gj - gprev_j, dx = x_j - xprev_j;
        a[0] = pr * pr; a[1] = dg * dg; a[2] = dg * dx; a[3] = dx * dx;
      }
      slice_sums(a, 4, 0);
      double b2[4] = {gval_part, 0.0, 0.0, 0.0};
      slice_sums(b2, 1, 4);
      if (t == 0) exch[5] = fs;
    }
    cluster_sync_all();                                   // barrier 2
    cluster_totals();
    const double gamma_prev = gamma;
    rule_step(O, totals[1], totals[2], totals[3], gamma, sigma, s0, s1);        // :341
    norm_res = sqrt(norm_sq_jl(totals[0]) + adapgm_dual_res_sq(gamma, gamma_prev, sigma));   // :348 (dual part: 0, or NaN -- phases.cuh)
    if (!(gamma == gamma) || !(norm_res == norm_res) || isinf(gamma)) flags |= ADAPROX_FLAG_NONFINITE;
    if (rank == 0 && t == 0 && W.rec != nullptr && it <= O.max_records) {
      adaprox_record rc;
      rc.it = it; rc.gamma = gamma; rc.sigma = sigma; rc.norm_res = norm_res;
      rc.f_x = 0.5 * norm_sq_jl(totals[5]);
      rc.g_x = want_obj ? prox_value_finish(P.g.kind, P.g.lambda, totals[4]) : NAN;
      rc.h_Ax = want_obj ? 0.0 : NAN;
      rc.f_evals = n_eval; rc.grad_f_evals = n_grad; rc.prox_g_evals = n_proxg; rc.prox_h_evals = 0;
      rc.A_evals = 0; rc.At_evals = 0;
      W.rec[it - 1] = rc;
    }
    if (it <= O.max_records) n_rec = it;
    if (norm_res <= O.tol) { converged = true; it_done = it; break; }           // :354-356 (uniform over the cluster)
    prox_and_broadcast(gj);
    n_proxg++;
    cluster_sync_all();                                   // barrier 3: x complete everywhere; exch / gp free again
  }

  if (own) W.xout[j] = x_j;                                // converged: the iterate whose gradient was just evaluated; maxit: the last prox
  if (rank == 0 && t == 0) {
    DResult r;
    r.iters = it_done;
    r.flags = flags | (converged ? ADAPROX_FLAG_CONVERGED : 0u);
    r.xbuf = 0;
    r.f_evals = n_eval; r.grad_f_evals = n_grad; r.prox_g_evals = n_proxg; r.prox_h_evals = 0;
    r.A_evals = 0; r.At_evals = 0; r.n_records = n_rec;
    r.final_gamma = gamma; r.final_sigma = sigma; r.final_norm_res = norm_res;
    *W.res = r;
  }
  cluster_sync_all();                                      // no CTA exits while a peer could still address its shared memory
}

}  // namespace adaprox
