// path.inl -- included at the end of api.cu: host side of the batched multi-lambda lasso path (path_gemm.cuh).
#include "path_gemm.cuh"

namespace adaprox {

static int path_setup_kernels(adaprox_ctx* h) {
  static bool done = false;
  if (done) return ADAPROX_OK;
  AP_CUDA(h, cudaFuncSetAttribute((const void*)k_path_gemm<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<4>::SmemBytes));
  AP_CUDA(h, cudaFuncSetAttribute((const void*)k_path_gemm<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<4>::SmemBytes));
  AP_CUDA(h, cudaFuncSetAttribute((const void*)k_path_gemm<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<1>::SmemBytes));
  AP_CUDA(h, cudaFuncSetAttribute((const void*)k_path_gemm<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<1>::SmemBytes));
  done = true;
  return ADAPROX_OK;
}

// G = A' R has only ceil(n / 128) * ceil(L / BN) output tiles; when that leaves most SMs idle (narrow batches) K = m is cut
// into slabs written to separate buffers and summed in slab order by k_path_step.
// mdim = rows of the output tile grid (n for G = A' R, m for R = A X), kdim = the contracted dimension
static int path_ksplit(adaprox_ctx* h, int64_t mdim, int64_t kdim, int64_t L) {
  const int bn = (L <= 64) ? GemmCfg<1>::BN : GemmCfg<4>::BN;
  const int64_t ctas = ((mdim + kGBM - 1) / kGBM) * ((L + bn - 1) / bn);
  const int64_t m = kdim;
  // makespan of ctas * ks equal work items on sm_count SMs, in units of one unsplit tile: ceil(ctas ks / SMs) / ks.
  // E.g. 128 tiles on 148 SMs: ks = 1 .. 7 all give 1.0 (20 SMs idle), ks = 8 gives 7/8.  Each slab costs one extra
  // L x n buffer that k_path_step reads, so only a clear gain (> 8 %) justifies splitting.
  int best = 1;
  double best_span = (double)((ctas + h->sm_count - 1) / h->sm_count);
  for (int ks = 2; ks <= 8; ++ks) {
    if (m / ks < 8 * kGBK) break;
    const double span = (double)((ctas * ks + h->sm_count - 1) / h->sm_count) / ks;
    if (span < 0.92 * best_span) { best = ks; best_span = span; }
  }
  return best;
}

// The batch is the N dimension of both contractions; narrow batches (L <= 64: e.g. 32 lambdas per rank when the path is
// split over 8 GPUs) use the 128 x 32 tile.
static void path_launch_gemm(adaprox_ctx* h, int mode, const PathGemmArgs& g) {
  const bool narrow = g.L <= 64;
  const int bn = narrow ? GemmCfg<1>::BN : GemmCfg<4>::BN;
  const int64_t Mdim = (mode == 1) ? g.m : g.n;
  const int ksp = (mode == 2) ? g.ksplit : g.ksplit1;
  dim3 grid((unsigned)((Mdim + kGBM - 1) / kGBM), (unsigned)((g.L + bn - 1) / bn), (unsigned)(ksp > 1 ? ksp : 1));
  if (mode == 1) {
    if (narrow) k_path_gemm<1, 1><<<grid, kGT, GemmCfg<1>::SmemBytes, h->stream>>>(g);
    else k_path_gemm<1, 4><<<grid, kGT, GemmCfg<4>::SmemBytes, h->stream>>>(g);
  } else {
    if (narrow) k_path_gemm<2, 1><<<grid, kGT, GemmCfg<1>::SmemBytes, h->stream>>>(g);
    else k_path_gemm<2, 4><<<grid, kGT, GemmCfg<4>::SmemBytes, h->stream>>>(g);
  }
  h->launches++;
}

// R = A X - b (+ fpart): one launch, or K-split partial products followed by the fix-up kernel
static void path_launch_r(adaprox_ctx* h, PathGemmArgs& g, double* RT, double* rslab) {
  if (g.ksplit1 > 1) {
    g.RT = rslab;
    path_launch_gemm(h, 1, g);
    g.RT = RT;
    dim3 grid((unsigned)((g.m + kRFixRows - 1) / kRFixRows), (unsigned)g.L);
    k_path_rfix<<<grid, 256, 0, h->stream>>>(g, rslab);
    h->launches++;
  } else {
    g.RT = RT;
    path_launch_gemm(h, 1, g);
  }
}

}  // namespace adaprox

using namespace adaprox;

extern "C" int adaprox_solve_lambda_path(adaprox_handle h, const adaprox_problem* p, const adaprox_options* o, int64_t L,
                                         const double* lambdas, const double* gamma0, const double* x0T, double* x_outT,
                                         int64_t* iters, double* norm_res, double* gamma_out, double* f_out,
                                         double* hist, int64_t hist_rows, adaprox_result* res) {
  if (!h || !p || !o || !lambdas || !x_outT || !res || L < 1) return fail(h, ADAPROX_ERR_INVALID, "solve_lambda_path: bad arguments");
  AP_CUDA(h, cudaSetDevice(h->device));
  int rc;
  if (o->solver != ADAPROX_S_ADAPTIVE_PROXGRAD)
    return fail(h, ADAPROX_ERR_UNSUPPORTED, "solve_lambda_path: adaptive_proxgrad / fixed_proxgrad only");
  if ((rc = validate_options(h, p, o))) return rc;
  DProblem P;
  HostMatrix *fm, *am;
  if ((rc = fill_problem(h, p, &P, &fm, &am))) return rc;
  if (P.f_kind != ADAPROX_F_LEAST_SQUARES || P.F.kind != MAT_DENSE)
    return fail(h, ADAPROX_ERR_UNSUPPORTED, "solve_lambda_path: the smooth term must be a dense least-squares term");
  if (fm && fm->sharded) return fail(h, ADAPROX_ERR_UNSUPPORTED, "solve_lambda_path: split the lambdas over the ranks, not the rows");
  for (int64_t j = 0; j < L; ++j)
    if (!(lambdas[j] >= 0.0)) return fail(h, ADAPROX_ERR_INVALID, "solve_lambda_path: lambdas must be >= 0");
  if ((rc = path_setup_kernels(h))) return rc;
  DOpts O{};
  fill_opts(o, &O);
  const int64_t nrec = hist ? std::min<int64_t>(hist_rows, O.maxit) : 0;
  O.max_records = nrec;
  const int64_t n = P.n, m = P.F.m;
  const int64_t ldx = round_up(n, 16), ldr = round_up(m, 16);
  const int ksplit = path_ksplit(h, n, m, L);             // G = A' R: tiles over n, contraction over m
  const int ksplit1 = path_ksplit(h, m, n, L);            // R = A X:  tiles over m, contraction over n
  const int64_t mtiles = ksplit1 > 1 ? (m + kRFixRows - 1) / kRFixRows : (m + kGBM - 1) / kGBM;
  size_t need = 7 * ws_size_doubles(L * ldx) + ws_size_doubles(L * ldr) + ws_size_doubles(mtiles * L) + ws_size_doubles((int64_t)ksplit * L * ldx) +
                (ksplit1 > 1 ? ws_size_doubles((int64_t)ksplit1 * L * ldr) : 0) +
                ws_size_doubles((L * (int64_t)sizeof(PathCol) + 7) / 8) + 3 * ws_size_doubles(std::max<int64_t>(nrec, 1) * L) + ws_size_doubles(1);
  if ((rc = ws_reset(h, need))) return rc;
  PathStepArgs sa{};
  sa.n = n; sa.ldx = ldx; sa.L = L; sa.mtiles = mtiles; sa.O = O;
  for (int k = 0; k < 3; ++k) sa.XT[k] = ws_doubles(h, L * ldx);
  for (int k = 0; k < 2; ++k) sa.GT[k] = ws_doubles(h, L * ldx);
  sa.VT = ws_doubles(h, L * ldx);
  sa.XoutT = ws_doubles(h, L * ldx);
  double* RT = ws_doubles(h, L * ldr);
  double* fpart = ws_doubles(h, mtiles * L);
  sa.fpart = fpart;
  double* gslab = ws_doubles(h, (int64_t)ksplit * L * ldx);
  double* rslab = ksplit1 > 1 ? ws_doubles(h, (int64_t)ksplit1 * L * ldr) : nullptr;
  sa.col = reinterpret_cast<PathCol*>(ws_doubles(h, (L * (int64_t)sizeof(PathCol) + 7) / 8));
  double* histd = ws_doubles(h, 3 * std::max<int64_t>(nrec, 1) * L);
  if (nrec > 0) { sa.gamma_hist = histd; sa.res_hist = histd + nrec * L; sa.obj_hist = histd + 2 * nrec * L; }
  sa.n_active = reinterpret_cast<int*>(ws_doubles(h, 1));

  // zero everything once: the padding of XT / RT is read by the tile loads and must stay zero
  AP_CUDA(h, cudaMemsetAsync(sa.XT[0], 0, (size_t)(7 * ws_size_doubles(L * ldx) + ws_size_doubles(L * ldr)), h->stream));
  if (nrec > 0) {
    std::vector<double> nanv((size_t)(3 * nrec * L), std::numeric_limits<double>::quiet_NaN());
    AP_CUDA(h, cudaMemcpyAsync(histd, nanv.data(), nanv.size() * 8, cudaMemcpyHostToDevice, h->stream));
    AP_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  if (x0T) AP_CUDA(h, cudaMemcpy2DAsync(sa.XT[0], (size_t)ldx * 8, x0T, (size_t)n * 8, (size_t)n * 8, (size_t)L, cudaMemcpyHostToDevice, h->stream));
  std::vector<PathCol> cols((size_t)L);
  for (int64_t j = 0; j < L; ++j) {
    PathCol& c = cols[(size_t)j];
    std::memset(&c, 0, sizeof(c));
    DOpts Oj = O;
    if (gamma0) Oj.gamma = gamma0[j];
    rule_init(Oj, c.gamma, c.sigma, c.s0, c.s1);                       // stepsize(rule) per column (:324)
    c.norm_res = INFINITY; c.lambda = lambdas[j];
  }
  AP_CUDA(h, cudaMemcpyAsync(sa.col, cols.data(), cols.size() * sizeof(PathCol), cudaMemcpyHostToDevice, h->stream));

  PathGemmArgs g{};
  g.A = P.F.a; g.m = m; g.n = n; g.lda = P.F.ld; g.b = P.fvec; g.L = L; g.ldx = ldx; g.RT = RT; g.ldr = ldr; g.fpart = fpart;
  g.ksplit = ksplit; g.gstride = L * ldx;
  g.ksplit1 = ksplit1; g.rstride = L * ldr;
  sa.ksplit = ksplit; sa.gstride = L * ldx; sa.Gslab = gslab;
  const int64_t launches0 = h->launches;
  AP_CUDA(h, cudaEventRecord(h->ev0, h->stream));
  int64_t evals = 0;
  int active = (int)L;
  for (int64_t it = 0; it <= O.maxit && active > 0; ++it) {
    g.XT = sa.XT[it % 3];
    g.GT = ksplit > 1 ? gslab : sa.GT[it & 1];
    path_launch_r(h, g, RT, rslab);                                    // R = A X - b, per-tile sums of r^2   (:336 value)
    g.RT = RT;
    path_launch_gemm(h, 2, g);                                         // G = A' R                            (:336 pullback)
    ++evals;
    AP_CUDA(h, cudaMemsetAsync(sa.n_active, 0, sizeof(int), h->stream));
    sa.it = it;
    k_path_step<<<(unsigned)L, kPT, 0, h->stream>>>(sa);
    h->launches++;
    AP_CUDA(h, cudaMemcpyAsync(&active, sa.n_active, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    AP_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  AP_CUDA(h, cudaEventRecord(h->ev1, h->stream));
  AP_CUDA(h, cudaGetLastError());
  AP_CUDA(h, cudaMemcpy2DAsync(x_outT, (size_t)n * 8, sa.XoutT, (size_t)ldx * 8, (size_t)n * 8, (size_t)L, cudaMemcpyDeviceToHost, h->stream));
  AP_CUDA(h, cudaMemcpyAsync(cols.data(), sa.col, cols.size() * sizeof(PathCol), cudaMemcpyDeviceToHost, h->stream));
  if (nrec > 0) AP_CUDA(h, cudaMemcpyAsync(hist, histd, (size_t)(3 * nrec * L) * 8, cudaMemcpyDeviceToHost, h->stream));
  AP_CUDA(h, cudaStreamSynchronize(h->stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, h->ev0, h->ev1);
  std::memset(res, 0, sizeof(*res));
  int64_t itmax = 0;
  unsigned all_conv = ADAPROX_FLAG_CONVERGED;
  for (int64_t j = 0; j < L; ++j) {
    const PathCol& c = cols[(size_t)j];
    if (iters) iters[j] = c.it_done;
    if (norm_res) norm_res[j] = c.norm_res;
    if (gamma_out) gamma_out[j] = c.gamma;
    if (f_out) f_out[j] = c.f_x;
    itmax = std::max<int64_t>(itmax, c.it_done);
    if (!(c.flags & ADAPROX_FLAG_CONVERGED)) all_conv = 0;
  }
  res->iters = itmax; res->flags = all_conv;
  res->f_evals = evals; res->grad_f_evals = evals; res->prox_g_evals = evals;      // batched oracle calls (each covers L columns)
  res->n_records = nrec;
  res->solve_ms = ms; res->kernel_launches = h->launches - launches0;
  res->matrix_passes = 2;
  return ADAPROX_OK;
}

extern "C" int adaprox_time_path_gemm(adaprox_handle h, adaprox_id mat, int64_t L, int which, int reps, double* ms_per_launch) {
  if (!h || !ms_per_launch || reps <= 0 || L < 1 || (which != 0 && which != 1)) return fail(h, ADAPROX_ERR_INVALID, "time_path_gemm: bad arguments");
  HostMatrix* hm;
  int rc = get_mat(h, mat, &hm);
  if (rc) return rc;
  if (hm->d.kind != MAT_DENSE) return fail(h, ADAPROX_ERR_UNSUPPORTED, "time_path_gemm: dense matrices only");
  AP_CUDA(h, cudaSetDevice(h->device));
  if ((rc = path_setup_kernels(h))) return rc;
  const DMat& M = hm->d;
  const int64_t ldx = round_up(M.n, 16), ldr = round_up(M.m, 16);
  const int ksplit = path_ksplit(h, M.n, M.m, L), ksplit1 = path_ksplit(h, M.m, M.n, L);
  const int64_t mtiles = ksplit1 > 1 ? (M.m + kRFixRows - 1) / kRFixRows : (M.m + kGBM - 1) / kGBM;
  size_t need = ws_size_doubles(L * ldx) + ws_size_doubles((int64_t)ksplit * L * ldx) + ws_size_doubles(L * ldr) + ws_size_doubles(mtiles * L) +
                ws_size_doubles(M.m) + (ksplit1 > 1 ? ws_size_doubles((int64_t)ksplit1 * L * ldr) : 0);
  if ((rc = ws_reset(h, need))) return rc;
  PathGemmArgs g{};
  double* XT = ws_doubles(h, L * ldx);
  g.GT = ws_doubles(h, (int64_t)ksplit * L * ldx);
  double* RT = ws_doubles(h, L * ldr);
  g.RT = RT;
  g.fpart = ws_doubles(h, mtiles * L);
  double* bz = ws_doubles(h, M.m);
  double* rslab = ksplit1 > 1 ? ws_doubles(h, (int64_t)ksplit1 * L * ldr) : nullptr;
  AP_CUDA(h, cudaMemsetAsync(XT, 0, need, h->stream));
  g.A = M.a; g.m = M.m; g.n = M.n; g.lda = M.ld; g.b = bz; g.L = L; g.XT = XT; g.ldx = ldx; g.ldr = ldr;
  g.ksplit = ksplit; g.gstride = L * ldx; g.ksplit1 = ksplit1; g.rstride = L * ldr;
  auto launch = [&]() { if (which == 0) path_launch_r(h, g, RT, rslab); else { g.RT = RT; path_launch_gemm(h, 2, g); } };
  launch();                                                            // warm-up
  AP_CUDA(h, cudaEventRecord(h->ev0, h->stream));
  for (int r = 0; r < reps; ++r) launch();
  AP_CUDA(h, cudaEventRecord(h->ev1, h->stream));
  AP_CUDA(h, cudaStreamSynchronize(h->stream));
  AP_CUDA(h, cudaGetLastError());
  float ms = 0.f;
  cudaEventElapsedTime(&ms, h->ev0, h->ev1);
  *ms_per_launch = (double)ms / reps;
  return ADAPROX_OK;
}
