// hessian.inl -- included at the end of api.cu.
//
// logistic_loss_grad_Hessian(X, y, w) of experiments/cubic_sparse_logreg/runme.jl:34-45: the setup step that builds the
// Cubic oracle's (Q, q) from a logistic-regression data set,
//     probs = sigm.(X w[1:end-1] .+ w[end]);  g = [X'(probs - y) / N ; mean(probs - y)];  sb = probs (1 - probs) / N
//     H = [X' R X   X' sb ; (X' sb)'   sum(sb)],   R = diagm(sb)
// g is the LogisticLoss gradient (the same kernels as eval_with_pullback); H is accumulated one output row per CTA with
// a FIXED summation order over the samples (no fp64 atomics, reruns are bit-identical).  A one-off setup: the data sets of
// the experiment have n <= a few hundred features, H is (n+1)^2 dense.

namespace adaprox {

// dense X (row-major [m][ld]): CTA a forms row a of H; thread b owns column b
__global__ void __launch_bounds__(256) k_logistic_hessian_dense(DMat X, const double* __restrict__ rvec, const double* __restrict__ yvec,
                                                                double invN, double* __restrict__ H, int64_t ldh) {
  const int64_t a = blockIdx.x, n = X.n;
  for (int64_t b0 = 0; b0 <= n; b0 += blockDim.x) {
    const int64_t b = b0 + threadIdx.x;
    double acc = 0.0;
    for (int64_t i = 0; i < X.m; ++i) {
      const double p = rvec[i] + yvec[i];                 // probs = (probs - y) + y
      const double sb = p * (1.0 - p) * invN;             // :40
      const double left = (a < n) ? X.a[i * X.ld + a] * sb : sb;          // (X' R)[a, i]; last row: R * ones
      const double right = (b < n) ? X.a[i * X.ld + (b < n ? b : 0)] : 1.0;
      acc = fma(left, right, acc);
    }
    if (b <= n) H[a * ldh + b] = acc;
  }
}

// CSR X: one warp per output row a, the row accumulated in shared memory; samples in the order of CSR(X') (ascending i)
__global__ void __launch_bounds__(32) k_logistic_hessian_csr(DMat X, const double* __restrict__ rvec, const double* __restrict__ yvec,
                                                             double invN, double* __restrict__ H, int64_t ldh) {
  extern __shared__ double s_row[];
  const int64_t a = blockIdx.x, n = X.n;
  const int lane = threadIdx.x;
  for (int64_t b = lane; b <= n; b += 32) s_row[b] = 0.0;
  __syncwarp();
  auto add_sample = [&](int64_t i, double left) {         // s_row += left * [X[i, :], 1]
    for (int64_t p = X.rowptr[i] + lane; p < X.rowptr[i + 1]; p += 32) s_row[X.colind[p]] = fma(left, X.vals[p], s_row[X.colind[p]]);
    if (lane == 0) s_row[n] += left;
    __syncwarp();
  };
  if (a < n) {
    for (int64_t k = X.t_rowptr[a]; k < X.t_rowptr[a + 1]; ++k) {
      const int64_t i = X.t_colind[k];
      const double p = rvec[i] + yvec[i];
      add_sample(i, X.t_vals[k] * (p * (1.0 - p) * invN));
    }
  } else {
    for (int64_t i = 0; i < X.m; ++i) {
      const double p = rvec[i] + yvec[i];
      add_sample(i, p * (1.0 - p) * invN);
    }
  }
  for (int64_t b = lane; b <= n; b += 32) H[a * ldh + b] = s_row[b];
}

}  // namespace adaprox

extern "C" int adaprox_logistic_grad_hessian(adaprox_handle h, adaprox_id X_mat, adaprox_id y_vec, const double* w,
                                             double* H_out, double* g_out) {
  using namespace adaprox;
  if (!h || !w || !H_out || !g_out) return fail(h, ADAPROX_ERR_INVALID, "logistic_grad_hessian: bad arguments");
  AP_CUDA(h, cudaSetDevice(h->device));
  HostMatrix* xm;
  int rc = get_mat(h, X_mat, &xm);
  if (rc) return rc;
  if (xm->sharded) return fail(h, ADAPROX_ERR_UNSUPPORTED, "logistic_grad_hessian: a row shard is not supported");
  adaprox_problem q{};
  q.f_kind = ADAPROX_F_LOGISTIC; q.f_mat = X_mat; q.f_vec = y_vec; q.n = xm->d.n + 1;
  DProblem P;
  HostMatrix *fm, *am;
  if ((rc = fill_problem(h, &q, &P, &fm, &am))) return rc;
  const int64_t n1 = P.n, m = P.F.m;
  if (P.F.kind == MAT_CSR && (size_t)n1 * 8 > 200 * 1024)
    return fail(h, ADAPROX_ERR_UNSUPPORTED, "logistic_grad_hessian: more than 25599 features (the Hessian row does not fit in shared memory; H would be > 5 GB)");
  if ((rc = ws_reset(h, 2 * ws_size_doubles(n1) + ws_size_doubles(std::max(m, n1)) + ws_size_doubles(1) + ws_size_doubles((int64_t)kMaxRed * h->grid) +
                            ws_size_doubles(4) + ws_size_doubles(n1 * n1)))) return rc;
  double* dx = ws_doubles(h, n1);
  double* dg = ws_doubles(h, n1);
  DWork W{};
  W.r = ws_doubles(h, std::max(m, n1));
  W.fu = ws_doubles(h, 1);
  W.red = ws_doubles(h, (int64_t)kMaxRed * h->grid);
  double* scal = ws_doubles(h, 4);
  double* dH = ws_doubles(h, n1 * n1);
  AP_CUDA(h, cudaMemcpyAsync(dx, w, (size_t)n1 * 8, cudaMemcpyHostToDevice, h->stream));
  OpArgs a{}; a.op = OP_EVALF; a.P = P; a.in = dx; a.out = dg; a.scal = scal; a.want_grad = 1;
  if ((rc = run_ops(h, a, W))) return rc;                  // leaves W.r = probs - y (sparse_logreg/runme.jl:36) and dg = g
  const double invN = 1.0 / (double)m;
  if (P.F.kind == MAT_DENSE) {
    k_logistic_hessian_dense<<<(unsigned)n1, 256, 0, h->stream>>>(P.F, W.r, P.fvec, invN, dH, n1);
  } else {
    AP_CUDA(h, cudaFuncSetAttribute((const void*)k_logistic_hessian_csr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(n1 * 8)));
    k_logistic_hessian_csr<<<(unsigned)n1, 32, (size_t)n1 * 8, h->stream>>>(P.F, W.r, P.fvec, invN, dH, n1);
  }
  h->launches++;
  AP_CUDA(h, cudaGetLastError());
  AP_CUDA(h, cudaMemcpyAsync(H_out, dH, (size_t)n1 * n1 * 8, cudaMemcpyDeviceToHost, h->stream));   // symmetric: row-major == column-major
  AP_CUDA(h, cudaMemcpyAsync(g_out, dg, (size_t)n1 * 8, cudaMemcpyDeviceToHost, h->stream));
  AP_CUDA(h, cudaStreamSynchronize(h->stream));
  return ADAPROX_OK;
}
