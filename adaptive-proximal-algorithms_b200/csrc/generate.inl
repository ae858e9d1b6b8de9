// generate.inl -- included at the end of api.cu.
//
// Device-side construction of the planted lasso instance of the reference's
// experiments/lasso/runme.jl:40-77 for one row shard.  The 65536 x 131072
// instance (68.7 GB) can never exist on the host, so the matrix is generated in
// place from the counter-based RNG (bit-identical to synth.py), and the steps
// that need global information (C'y*, A x*, opnorm) run through the same GEMV
// kernels as the solver, all-reduced across ranks when a communicator is attached.

extern "C" int adaprox_generate_planted_lasso(adaprox_handle h, int64_t m, int64_t n, int64_t row0, int64_t rows,
                                              double pfactor, uint64_t seed, double lam, double rho, int32_t power_iters,
                                              adaprox_id* A_out, adaprox_id* b_out, double* x_star, double* optimum, double* Lf) {
  using namespace adaprox;
  if (!h || !A_out || !b_out || m <= 0 || n <= 0 || row0 < 0 || rows <= 0 || row0 + rows > m || !(pfactor > 0))
    return fail(h, ADAPROX_ERR_INVALID, "generate_planted_lasso: bad arguments");
  const bool sharded = (rows != m);
  if (sharded && !h->comm) return fail(h, ADAPROX_ERR_COMM, "generate_planted_lasso: a row shard needs a communicator");
  AP_CUDA(h, cudaSetDevice(h->device));
  HostMatrix hm;
  int rc = alloc_dense(h, rows, n, hm);
  if (rc) { free_matrix(hm); return rc; }
  hm.m_global = m; hm.row0 = row0; hm.sharded = sharded;
  DMat& M = hm.d;
  double* a = const_cast<double*>(M.a);
  auto bail = [&](int code) { free_matrix(hm); return code; };

  // y_star = rand(m); y_star ./= norm(y_star)   (:48-49) -- every rank builds the full vector
  std::vector<double> ystar((size_t)m);
  {
    const uint64_t key = stream_key(seed, 0);
    double ss = 0.0;
    for (int64_t i = 0; i < m; ++i) { ystar[i] = uniform01(key, (uint64_t)i); ss += ystar[i] * ystar[i]; }
    const double nrm = std::sqrt(ss);
    for (int64_t i = 0; i < m; ++i) ystar[i] /= nrm;
  }
  // C = rand(m, n) .* 2 .- 1   (:50), rows [row0, row0 + rows)
  k_fill_uniform_pm1<<<h->sm_count * 8, 256, 0, h->stream>>>(a, rows, n, M.ld, row0, stream_key(seed, 1));
  h->launches++;

  double *d_ys = nullptr, *d_n = nullptr, *d_n2 = nullptr, *d_m = nullptr;
  auto cleanup = [&]() { cudaFree(d_ys); cudaFree(d_n); cudaFree(d_n2); cudaFree(d_m); };
#define GEN_CUDA(call)                                                                                   \
  do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(); free_matrix(hm);                    \
       return fail(h, ADAPROX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); } } while (0)
#define GEN_RC(call) do { int r__ = (call); if (r__) { cleanup(); return bail(r__); } } while (0)
  GEN_CUDA(cudaMalloc(&d_ys, (size_t)rows * 8));
  GEN_CUDA(cudaMalloc(&d_n, (size_t)n * 8));
  GEN_CUDA(cudaMalloc(&d_n2, (size_t)n * 8));
  GEN_CUDA(cudaMalloc(&d_m, (size_t)rows * 8));
  GEN_CUDA(cudaMemcpyAsync(d_ys, ystar.data() + row0, (size_t)rows * 8, cudaMemcpyHostToDevice, h->stream));

  // CTy = abs.(C' * y_star)   (:52)
  GEN_RC(op_amul_dev(h, M, d_ys, d_n));
  if (sharded) GEN_RC(comm_allreduce_sum(h, d_n, n));
  std::vector<double> cty((size_t)n), sgn((size_t)n);
  GEN_CUDA(cudaMemcpyAsync(cty.data(), d_n, (size_t)n * 8, cudaMemcpyDeviceToHost, h->stream));
  GEN_CUDA(cudaStreamSynchronize(h->stream));
  for (int64_t j = 0; j < n; ++j) { sgn[j] = cty[j] > 0 ? 1.0 : (cty[j] < 0 ? -1.0 : 0.0); cty[j] = std::fabs(cty[j]); }
  // perm = sortperm(CTy, rev = true)   (:53) -- stable
  std::vector<int64_t> perm((size_t)n);
  std::iota(perm.begin(), perm.end(), 0);
  std::stable_sort(perm.begin(), perm.end(), [&](int64_t x, int64_t y) { return cty[x] > cty[y]; });
  // alpha (:55-68) and x_star (:71-75)
  const double p = (double)n / pfactor;
  const uint64_t key_alpha = stream_key(seed, 2), key_x = stream_key(seed, 3);
  std::vector<double> alpha((size_t)n, 0.0), xs((size_t)n, 0.0);
  double l1 = 0.0;
  for (int64_t k = 0; k < n; ++k) {
    const int64_t j = perm[k];
    if ((double)(k + 1) <= p) {
      alpha[j] = lam / cty[j];
      // sign(dot(A[:, j], y_star)) = sign(alpha_j * (C'y*)_j) = sign((C'y*)_j)
      xs[j] = uniform01(key_x, (uint64_t)j) * rho / std::sqrt(p) * sgn[j];
      l1 += std::fabs(xs[j]);
    } else {
      const double temp = cty[j];
      alpha[j] = (temp < 0.1 * lam) ? lam : lam * uniform01(key_alpha, (uint64_t)j) / temp;
    }
  }
  // A = C * diagm(alpha)   (:69)
  GEN_CUDA(cudaMemcpyAsync(d_n, alpha.data(), (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
  k_scale_columns<<<h->sm_count * 8, 256, 0, h->stream>>>(a, rows, n, M.ld, d_n);
  h->launches++;
  // b = A * x_star + y_star   (:76)
  GEN_CUDA(cudaMemcpyAsync(d_n2, xs.data(), (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
  GEN_RC(op_mul_dev(h, M, d_n2, d_m));
  k_axpby<<<h->sm_count * 2, 256, 0, h->stream>>>(rows, 1.0, d_ys, 1.0, d_m);      // d_m = y_star + A x_star
  h->launches++;
  // Lf = opnorm(A)^2   (:81) by power iteration on A'A from the normalised all-ones vector
  double lf = 0.0;
  if (power_iters > 0) {
    std::vector<double> v((size_t)n, 1.0 / std::sqrt((double)n));
    GEN_CUDA(cudaMemcpyAsync(d_n, v.data(), (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
    double* d_tmp = nullptr;
    GEN_CUDA(cudaMalloc(&d_tmp, (size_t)rows * 8));
    for (int it = 0; it < power_iters; ++it) {
      int r1 = op_mul_dev(h, M, d_n, d_tmp);
      int r2 = r1 ? r1 : op_amul_dev(h, M, d_tmp, d_n2);
      if (!r2 && sharded) r2 = comm_allreduce_sum(h, d_n2, n);
      cudaError_t e = r2 ? cudaSuccess : cudaMemcpyAsync(v.data(), d_n2, (size_t)n * 8, cudaMemcpyDeviceToHost, h->stream);
      if (!r2 && e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
      if (r2 || e != cudaSuccess) { cudaFree(d_tmp); cleanup(); free_matrix(hm); return r2 ? r2 : fail(h, ADAPROX_ERR_CUDA, cudaGetErrorString(e)); }
      double ss = 0.0;
      for (int64_t j = 0; j < n; ++j) ss += v[j] * v[j];
      lf = std::sqrt(ss);
      for (int64_t j = 0; j < n; ++j) v[j] /= lf;
      e = cudaMemcpyAsync(d_n, v.data(), (size_t)n * 8, cudaMemcpyHostToDevice, h->stream);
      if (e != cudaSuccess) { cudaFree(d_tmp); cleanup(); free_matrix(hm); return fail(h, ADAPROX_ERR_CUDA, cudaGetErrorString(e)); }
    }
    GEN_CUDA(cudaStreamSynchronize(h->stream));
    cudaFree(d_tmp);
  }
  GEN_CUDA(cudaStreamSynchronize(h->stream));
  // hand b over as a library vector
  HostVector hv;
  hv.len = rows; hv.p = d_m; d_m = nullptr;
  const int64_t bid = h->next_id++;
  h->vecs[bid] = hv;
  const int64_t aid = h->next_id++;
  h->mats[aid] = hm;
  cleanup();
#undef GEN_CUDA
#undef GEN_RC
  if (x_star) std::memcpy(x_star, xs.data(), (size_t)n * 8);
  if (optimum) *optimum = 1.0 / 2.0 + lam * l1;      // norm(y_star) / 2 + lam * norm(x_star, 1), ||y*|| = 1  (:77)
  if (Lf) *Lf = lf;
  *A_out = aid; *b_out = bid;
  return ADAPROX_OK;
}
