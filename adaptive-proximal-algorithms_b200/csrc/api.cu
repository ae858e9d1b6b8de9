// api.cu -- the C ABI of libadaprox_cuda.so (include/adaprox.h).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include "context.hpp"
#include "ops.cuh"
#include "solver_pd.cuh"
#include "solver_pg.cuh"
#include "solver_mp.cuh"
#include "solver_fused.cuh"
#include "solver_resident.cuh"
#include "solver_gridres.cuh"
#include "comm.hpp"

using namespace adaprox;

// ---------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------
static int64_t round_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

static int ws_reset(adaprox_ctx* h, size_t need) {
  if (need > h->ws_bytes) {
    if (h->ws) cudaFree(h->ws);
    h->ws = nullptr; h->ws_bytes = 0;
    size_t want = need + (need >> 3) + (1u << 20);
    cudaError_t e = cudaMalloc(&h->ws, want);
    if (e != cudaSuccess) return fail(h, ADAPROX_ERR_NOMEM, std::string("workspace cudaMalloc: ") + cudaGetErrorString(e));
    h->ws_bytes = want;
  }
  h->ws_used = 0;
  return ADAPROX_OK;
}
static double* ws_doubles(adaprox_ctx* h, int64_t count) {
  size_t bytes = (size_t)round_up(std::max<int64_t>(count, 1) * 8, 256);
  double* p = reinterpret_cast<double*>(h->ws + h->ws_used);
  h->ws_used += bytes;
  return p;
}
static size_t ws_size_doubles(int64_t count) { return (size_t)round_up(std::max<int64_t>(count, 1) * 8, 256); }

static int get_mat(adaprox_ctx* h, adaprox_id id, HostMatrix** out) {
  auto it = h->mats.find(id);
  if (it == h->mats.end()) return fail(h, ADAPROX_ERR_INVALID, "unknown matrix id " + std::to_string(id));
  *out = &it->second;
  return ADAPROX_OK;
}
static int get_vec(adaprox_ctx* h, adaprox_id id, int64_t min_len, const double** out) {
  if (id == 0) { *out = nullptr; return ADAPROX_OK; }
  auto it = h->vecs.find(id);
  if (it == h->vecs.end()) return fail(h, ADAPROX_ERR_INVALID, "unknown vector id " + std::to_string(id));
  if (it->second.len < min_len) return fail(h, ADAPROX_ERR_INVALID, "vector " + std::to_string(id) + " is too short");
  *out = it->second.p;
  return ADAPROX_OK;
}

template <typename K>
static int coop_launch(adaprox_ctx* h, K kernel, void** args, int grid = 0) {
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)kernel, dim3(grid > 0 ? grid : h->grid), dim3(kThreads), args, kRingBytes, h->stream);
  if (e != cudaSuccess) return fail(h, ADAPROX_ERR_CUDA, std::string("cooperative launch: ") + cudaGetErrorString(e));
  h->launches++;
  return ADAPROX_OK;
}

static constexpr double kL2KeepMB = 0.0;

// choose the work partition of a dense matrix for a grid of G CTAs
static void plan_dense(DMat& d, int G) {
  d.nchunks = (int)((d.n + kChunk - 1) / kChunk);
  d.npad = (int64_t)d.nchunks * kChunk;
  int rb = 64;
  while (rb > 8 && (int64_t)d.nchunks * ((d.m + rb - 1) / rb) < 16LL * G) rb >>= 1;
  d.rb = rb;
  d.nrb = (d.m + rb - 1) / rb;
  const char* e = std::getenv("ADAPROX_GEMV");       // "ldg": register-staged loads; default: bulk-copy ring
  d.path = (e && std::strcmp(e, "ldg") == 0) ? 0 : 1;
  // L2 budget the end of a sweep may pin for the start of the next (gemv.cuh, gemv_n_phase): MB over the whole grid
  const char* k = std::getenv("ADAPROX_L2_KEEP_MB");
  const double mb = k ? std::atof(k) : kL2KeepMB;
  const int64_t tile = std::min<int64_t>(d.ld, kChunk) * 8;
  d.keep = (mb < 0.0) ? -1 : (int64_t)(mb * 1048576.0 / G / (double)tile);
  d.alternate = std::getenv("ADAPROX_SWEEP_ONE_WAY") ? 0 : 1;
}

static int alloc_dense(adaprox_ctx* h, int64_t m, int64_t n, HostMatrix& hm) {
  DMat& d = hm.d;
  d = DMat{};
  d.kind = MAT_DENSE; d.m = m; d.n = n; d.ld = round_up(n, 16);
  plan_dense(d, h->grid);
  double *a = nullptr, *zp = nullptr, *gp = nullptr;
  AP_CUDA(h, cudaMalloc(&a, (size_t)m * d.ld * 8));
  hm.allocs.push_back(a);
  AP_CUDA(h, cudaMalloc(&zp, (size_t)d.nchunks * m * 8));
  hm.allocs.push_back(zp);
  AP_CUDA(h, cudaMalloc(&gp, (size_t)h->grid * d.npad * 8));
  hm.allocs.push_back(gp);
  AP_CUDA(h, cudaMemsetAsync(gp, 0, (size_t)h->grid * d.npad * 8, h->stream));
  d.a = a; d.zpart = zp; d.gpart = gp;
  hm.m_global = m; hm.row0 = 0; hm.sharded = false;
  return ADAPROX_OK;
}

static void free_matrix(HostMatrix& hm) {
  for (void* p : hm.allocs) cudaFree(p);
  hm.allocs.clear();
}

// device-pointer operator calls (shared by the public ops and the generator)
static int run_ops(adaprox_ctx* h, OpArgs& a, DWork& W) {
  void* args[] = {&a, &W};
  return coop_launch(h, k_ops, args);
}

static int op_mul_dev(adaprox_ctx* h, const DMat& M, const double* x_dev, double* out_dev) {
  OpArgs a{}; a.op = OP_MUL; a.M = M; a.in = x_dev; a.out = out_dev;
  DWork W{};
  return run_ops(h, a, W);
}
static int op_amul_dev(adaprox_ctx* h, const DMat& M, const double* y_dev, double* out_dev) {
  OpArgs a{}; a.op = OP_AMUL; a.M = M; a.in = y_dev; a.out = out_dev;
  DWork W{};
  return run_ops(h, a, W);
}

static int fill_prox(adaprox_ctx* h, const adaprox_prox& p, int64_t len, DProx* out) {
  DProx d{};
  d.kind = p.kind; d.conjugate = p.conjugate; d.lambda = p.lambda; d.lo = p.lo; d.hi = p.hi;
  if (p.kind < ADAPROX_P_ZERO || p.kind > ADAPROX_P_IND_BOX) return fail(h, ADAPROX_ERR_INVALID, "unknown prox kind");
  int rc;
  if ((rc = get_vec(h, p.lo_vec, len, &d.lo_vec))) return rc;
  if ((rc = get_vec(h, p.hi_vec, len, &d.hi_vec))) return rc;
  if ((rc = get_vec(h, p.shift, len, &d.shift))) return rc;
  *out = d;
  return ADAPROX_OK;
}

static int fill_problem(adaprox_ctx* h, const adaprox_problem* p, DProblem* out, HostMatrix** fmat, HostMatrix** amat) {
  DProblem P{};
  P.f_kind = p->f_kind; P.f_ipar = p->f_ipar; P.f_c = p->f_c; P.n = p->n; P.md = p->m_dual;
  *fmat = nullptr; *amat = nullptr;
  int rc;
  if (p->n <= 0) return fail(h, ADAPROX_ERR_INVALID, "problem.n must be positive");
  int64_t vec_len = 0;
  switch (p->f_kind) {
    case ADAPROX_F_ZERO: break;
    case ADAPROX_F_LEAST_SQUARES:
    case ADAPROX_F_LOGISTIC:
    case ADAPROX_F_QUADRATIC:
    case ADAPROX_F_QUADRATIC_GRAM:
    case ADAPROX_F_CUBIC: {
      if ((rc = get_mat(h, p->f_mat, fmat))) return rc;
      P.F = (*fmat)->d;
      P.F.slot = 0;
      const bool gram = (p->f_kind == ADAPROX_F_QUADRATIC_GRAM);      // F = Z (n x d), any d
      const int64_t ncols = (p->f_kind == ADAPROX_F_LOGISTIC) ? p->n - 1 : p->n;
      if (gram && P.F.kind != MAT_DENSE) return fail(h, ADAPROX_ERR_UNSUPPORTED, "QuadraticGram: Z must be a dense matrix");
      if (!gram && P.F.n != ncols) return fail(h, ADAPROX_ERR_INVALID, "f matrix has " + std::to_string(P.F.n) + " columns, expected " + std::to_string(ncols));
      if ((p->f_kind == ADAPROX_F_QUADRATIC || p->f_kind == ADAPROX_F_CUBIC) && (*fmat)->m_global != p->n)
        return fail(h, ADAPROX_ERR_INVALID, "Q must be square");
      if (gram && (*fmat)->m_global != p->n) return fail(h, ADAPROX_ERR_INVALID, "QuadraticGram: Z must have n rows");
      if (p->f_kind == ADAPROX_F_CUBIC && (*fmat)->sharded) return fail(h, ADAPROX_ERR_UNSUPPORTED, "Cubic: a row-sharded Q is not supported");
      if (p->f_kind == ADAPROX_F_QUADRATIC || gram) P.f_row0 = (*fmat)->sharded ? (*fmat)->row0 : 0;
      vec_len = (p->f_kind == ADAPROX_F_QUADRATIC || gram || p->f_kind == ADAPROX_F_CUBIC) ? p->n : P.F.m;
      P.f_N = (double)((*fmat)->m_global);
      if (p->f_vec == 0) return fail(h, ADAPROX_ERR_INVALID, "f_vec (b / y / q) is required");
    } break;
    case ADAPROX_F_WORST_QUADRATIC:
      if (p->f_ipar < 2 || p->f_ipar > p->n) return fail(h, ADAPROX_ERR_INVALID, "WorstQuadratic needs 2 <= k <= n");
      break;
    case ADAPROX_F_SIMPLE2D:
      if (p->n != 2) return fail(h, ADAPROX_ERR_INVALID, "Simple2D needs n = 2");
      break;
    default:
      return fail(h, ADAPROX_ERR_UNSUPPORTED, "eval_with_pullback not defined for f_kind " + std::to_string(p->f_kind));
  }
  if ((rc = get_vec(h, p->f_vec, vec_len, &P.fvec))) return rc;
  if ((rc = fill_prox(h, p->g, p->n, &P.g))) return rc;
  if (p->A_mat != 0) {
    if ((rc = get_mat(h, p->A_mat, amat))) return rc;
    if (p->A_mat == p->f_mat)     // the partial buffers (zpart / gpart) belong to the matrix: F'r and A'y would share them in one phase
      return fail(h, ADAPROX_ERR_UNSUPPORTED, "f and A refer to the same device matrix: upload it a second time for A");
    P.A = (*amat)->d;
    P.A.slot = 1;
    if (P.A.n != p->n) return fail(h, ADAPROX_ERR_INVALID, "A has the wrong number of columns");
    if (P.A.m != p->m_dual) return fail(h, ADAPROX_ERR_INVALID, "A has the wrong number of rows (m_dual)");
    if ((rc = fill_prox(h, p->h, p->m_dual, &P.h))) return rc;
  } else {
    P.A.kind = MAT_NONE;
    P.h = DProx{};
  }
  *out = P;
  return ADAPROX_OK;
}

// ---------------------------------------------------------------------------
// life cycle
// ---------------------------------------------------------------------------
extern "C" int adaprox_version(void) { return 100; }

extern "C" int adaprox_create(adaprox_handle* out, int device) {
  if (!out) return ADAPROX_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return ADAPROX_ERR_CUDA;   // no CPU fallback
  if (device < 0 || device >= ndev) return ADAPROX_ERR_INVALID;
  if (cudaSetDevice(device) != cudaSuccess) return ADAPROX_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return ADAPROX_ERR_CUDA;
  if (!prop.cooperativeLaunch) return ADAPROX_ERR_UNSUPPORTED;
  adaprox_ctx* h = new adaprox_ctx();
  h->device = device; h->sm_count = prop.multiProcessorCount; h->cc_major = prop.major; h->cc_minor = prop.minor;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return ADAPROX_ERR_CUDA; }
  cudaEventCreate(&h->ev0); cudaEventCreate(&h->ev1);
  if (cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking) != cudaSuccess) { delete h; return ADAPROX_ERR_CUDA; }
  cudaEventCreateWithFlags(&h->ev_h, cudaEventDisableTiming);
  if (cudaHostAlloc((void**)&h->resident_host, 64, cudaHostAllocMapped) == cudaSuccess) {
    *h->resident_host = 0ull;
    if (cudaHostGetDevicePointer((void**)&h->resident_dev, h->resident_host, 0) != cudaSuccess) h->resident_dev = nullptr;
  } else {
    cudaGetLastError();
    h->resident_host = nullptr;
  }
  // resident CTAs per SM: the minimum over the persistent kernels
  int per_sm = 2, nb = 0;
  const void* kernels[] = {(const void*)k_primal_dual<false>, (const void*)k_primal_dual<true>, (const void*)k_ops,
                           (const void*)k_proxgrad_family, (const void*)k_gemv_pass, (const void*)k_malitsky_pock};
  for (const void* k : kernels) {
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kRingBytes) != cudaSuccess) {
      delete h; return ADAPROX_ERR_CUDA;
    }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, kThreads, kRingBytes) != cudaSuccess || nb < 1) {
      delete h; return ADAPROX_ERR_CUDA;
    }
    per_sm = std::min(per_sm, nb);
  }
  if (comm_setup_kernels() != 0) { delete h; return ADAPROX_ERR_CUDA; }
  h->grid = per_sm * h->sm_count;
  *out = h;
  return ADAPROX_OK;
}

extern "C" int adaprox_destroy(adaprox_handle h) {
  if (!h) return ADAPROX_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  comm_destroy(h);
  for (auto& kv : h->mats) free_matrix(kv.second);
  for (auto& kv : h->vecs) cudaFree(kv.second.p);
  if (h->ws) cudaFree(h->ws);
  cudaEventDestroy(h->ev0); cudaEventDestroy(h->ev1);
  if (h->ev_h) cudaEventDestroy(h->ev_h);
  if (h->resident_host) cudaFreeHost(h->resident_host);
  if (h->stream2) { cudaStreamSynchronize(h->stream2); cudaStreamDestroy(h->stream2); }
  cudaStreamDestroy(h->stream);
  delete h;
  return ADAPROX_OK;
}

extern "C" const char* adaprox_last_error(adaprox_handle h) { return h ? h->err.c_str() : "null handle"; }

extern "C" int adaprox_device_info(adaprox_handle h, int* sm_count, int* cc_major, int* cc_minor, int64_t* free_bytes) {
  if (!h) return ADAPROX_ERR_INVALID;
  if (sm_count) *sm_count = h->sm_count;
  if (cc_major) *cc_major = h->cc_major;
  if (cc_minor) *cc_minor = h->cc_minor;
  if (free_bytes) { size_t f = 0, t = 0; AP_CUDA(h, cudaMemGetInfo(&f, &t)); *free_bytes = (int64_t)f; }
  return ADAPROX_OK;
}

// ---------------------------------------------------------------------------
// device-resident data
// ---------------------------------------------------------------------------
extern "C" int adaprox_matrix_upload_colmajor(adaprox_handle h, const double* A, int64_t m, int64_t n, int64_t lda, adaprox_id* out) {
  if (!h || !A || !out || m <= 0 || n <= 0 || lda < m) return fail(h, ADAPROX_ERR_INVALID, "matrix_upload_colmajor: bad arguments");
  AP_CUDA(h, cudaSetDevice(h->device));
  HostMatrix hm;
  int rc = alloc_dense(h, m, n, hm);
  if (rc) { free_matrix(hm); return rc; }
  double* stage = nullptr;
  cudaError_t e = cudaMalloc(&stage, (size_t)lda * n * 8);
  if (e == cudaSuccess) e = cudaMemcpyAsync(stage, A, (size_t)lda * n * 8, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) {
    dim3 grid((unsigned)((m + 31) / 32), (unsigned)((hm.d.ld + 31) / 32)), block(32, 8);
    k_colmajor_to_rowmajor<<<grid, block, 0, h->stream>>>(stage, m, n, lda, const_cast<double*>(hm.d.a), hm.d.ld);
    h->launches++;
    e = cudaStreamSynchronize(h->stream);
  }
  if (stage) cudaFree(stage);
  if (e != cudaSuccess) { free_matrix(hm); return fail(h, ADAPROX_ERR_CUDA, std::string("matrix upload: ") + cudaGetErrorString(e)); }
  const int64_t id = h->next_id++;
  h->mats[id] = hm;
  *out = id;
  return ADAPROX_OK;
}

extern "C" int adaprox_matrix_upload_rowmajor(adaprox_handle h, const double* A, int64_t m, int64_t n, int64_t lda, adaprox_id* out) {
  if (!h || !A || !out || m <= 0 || n <= 0 || lda < n) return fail(h, ADAPROX_ERR_INVALID, "matrix_upload_rowmajor: bad arguments");
  AP_CUDA(h, cudaSetDevice(h->device));
  HostMatrix hm;
  int rc = alloc_dense(h, m, n, hm);
  if (rc) { free_matrix(hm); return rc; }
  cudaError_t e = cudaMemsetAsync(const_cast<double*>(hm.d.a), 0, (size_t)m * hm.d.ld * 8, h->stream);
  if (e == cudaSuccess)
    e = cudaMemcpy2DAsync(const_cast<double*>(hm.d.a), (size_t)hm.d.ld * 8, A, (size_t)lda * 8, (size_t)n * 8, (size_t)m,
                          cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) { free_matrix(hm); return fail(h, ADAPROX_ERR_CUDA, std::string("matrix upload: ") + cudaGetErrorString(e)); }
  const int64_t id = h->next_id++;
  h->mats[id] = hm;
  *out = id;
  return ADAPROX_OK;
}

extern "C" int adaprox_matrix_upload_csr(adaprox_handle h, int64_t m, int64_t n, int64_t nnz, const int64_t* rowptr,
                                         const int32_t* colind, const double* vals, adaprox_id* out) {
  if (!h || !rowptr || !out || m <= 0 || n <= 0 || nnz < 0 || (nnz > 0 && (!colind || !vals)))
    return fail(h, ADAPROX_ERR_INVALID, "matrix_upload_csr: bad arguments");
  if (rowptr[0] != 0 || rowptr[m] != nnz) return fail(h, ADAPROX_ERR_INVALID, "matrix_upload_csr: rowptr does not span nnz");
  for (int64_t i = 0; i < m; ++i)
    if (rowptr[i] > rowptr[i + 1]) return fail(h, ADAPROX_ERR_INVALID, "matrix_upload_csr: rowptr is not monotone");
  for (int64_t k = 0; k < nnz; ++k)
    if (colind[k] < 0 || colind[k] >= n) return fail(h, ADAPROX_ERR_INVALID, "matrix_upload_csr: column index out of range");
  AP_CUDA(h, cudaSetDevice(h->device));
  // CSR of the transpose (counting sort by column; rows stay sorted inside each column)
  std::vector<int64_t> tptr(n + 1, 0);
  for (int64_t k = 0; k < nnz; ++k) tptr[colind[k] + 1]++;
  for (int64_t j = 0; j < n; ++j) tptr[j + 1] += tptr[j];
  std::vector<int32_t> tind(std::max<int64_t>(nnz, 1));
  std::vector<double> tval(std::max<int64_t>(nnz, 1));
  {
    std::vector<int64_t> cur(tptr.begin(), tptr.end() - 1);
    for (int64_t i = 0; i < m; ++i)
      for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
        const int64_t pos = cur[colind[k]]++;
        tind[pos] = (int32_t)i; tval[pos] = vals[k];
      }
  }
  HostMatrix hm;
  DMat& d = hm.d;
  d.kind = MAT_CSR; d.m = m; d.n = n; d.ld = 0; d.nnz = nnz; d.nchunks = 1; d.rb = 0; d.nrb = 0; d.npad = n;
  d.lpr_n = csr_lanes_per_row(nnz, m); d.lpr_t = csr_lanes_per_row(nnz, n);
  if (const char* e = std::getenv("ADAPROX_CSR_LPR")) { const int v = std::atoi(e); if (v == 4 || v == 8 || v == 16 || v == 32) d.lpr_n = d.lpr_t = v; }
  auto up = [&](const void* src, size_t bytes, void** dst) -> cudaError_t {
    cudaError_t e = cudaMalloc(dst, std::max<size_t>(bytes, 8));
    if (e != cudaSuccess) return e;
    hm.allocs.push_back(*dst);
    return bytes ? cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) : cudaSuccess;
  };
  void *p_rp, *p_ci, *p_v, *p_trp, *p_tci, *p_tv, *p_zp, *p_gp;
  cudaError_t e = up(rowptr, (size_t)(m + 1) * 8, &p_rp);
  if (e == cudaSuccess) e = up(colind, (size_t)nnz * 4, &p_ci);
  if (e == cudaSuccess) e = up(vals, (size_t)nnz * 8, &p_v);
  if (e == cudaSuccess) e = up(tptr.data(), (size_t)(n + 1) * 8, &p_trp);
  if (e == cudaSuccess) e = up(tind.data(), (size_t)nnz * 4, &p_tci);
  if (e == cudaSuccess) e = up(tval.data(), (size_t)nnz * 8, &p_tv);
  if (e == cudaSuccess) { e = cudaMalloc(&p_zp, (size_t)m * 8); if (e == cudaSuccess) hm.allocs.push_back(p_zp); }
  if (e == cudaSuccess) { e = cudaMalloc(&p_gp, (size_t)n * 8); if (e == cudaSuccess) hm.allocs.push_back(p_gp); }
  if (e != cudaSuccess) { free_matrix(hm); return fail(h, ADAPROX_ERR_CUDA, std::string("csr upload: ") + cudaGetErrorString(e)); }
  d.rowptr = (const int64_t*)p_rp; d.colind = (const int*)p_ci; d.vals = (const double*)p_v;
  d.t_rowptr = (const int64_t*)p_trp; d.t_colind = (const int*)p_tci; d.t_vals = (const double*)p_tv;
  d.zpart = (double*)p_zp; d.gpart = (double*)p_gp;
  hm.m_global = m;
  const int64_t id = h->next_id++;
  h->mats[id] = hm;
  *out = id;
  return ADAPROX_OK;
}

extern "C" int adaprox_matrix_free(adaprox_handle h, adaprox_id mat) {
  if (!h) return ADAPROX_ERR_INVALID;
  auto it = h->mats.find(mat);
  if (it == h->mats.end()) return fail(h, ADAPROX_ERR_INVALID, "matrix_free: unknown id");
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  free_matrix(it->second);
  h->mats.erase(it);
  return ADAPROX_OK;
}

extern "C" int adaprox_matrix_shape(adaprox_handle h, adaprox_id mat, int64_t* m, int64_t* n, int64_t* nnz) {
  if (!h) return ADAPROX_ERR_INVALID;
  HostMatrix* hm;
  int rc = get_mat(h, mat, &hm);
  if (rc) return rc;
  if (m) *m = hm->d.m;
  if (n) *n = hm->d.n;
  if (nnz) *nnz = hm->d.kind == MAT_CSR ? hm->d.nnz : hm->d.m * hm->d.n;
  return ADAPROX_OK;
}

extern "C" int adaprox_vector_upload(adaprox_handle h, const double* v, int64_t len, adaprox_id* out) {
  if (!h || !v || !out || len <= 0) return fail(h, ADAPROX_ERR_INVALID, "vector_upload: bad arguments");
  AP_CUDA(h, cudaSetDevice(h->device));
  HostVector hv;
  hv.len = len;
  AP_CUDA(h, cudaMalloc(&hv.p, (size_t)len * 8));
  cudaError_t e = cudaMemcpy(hv.p, v, (size_t)len * 8, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(hv.p); return fail(h, ADAPROX_ERR_CUDA, cudaGetErrorString(e)); }
  const int64_t id = h->next_id++;
  h->vecs[id] = hv;
  *out = id;
  return ADAPROX_OK;
}

extern "C" int adaprox_vector_download(adaprox_handle h, adaprox_id vec, double* out, int64_t len) {
  if (!h || !out) return fail(h, ADAPROX_ERR_INVALID, "vector_download: bad arguments");
  auto it = h->vecs.find(vec);
  if (it == h->vecs.end() || it->second.len < len) return fail(h, ADAPROX_ERR_INVALID, "vector_download: unknown id or too short");
  AP_CUDA(h, cudaSetDevice(h->device));
  AP_CUDA(h, cudaStreamSynchronize(h->stream));
  AP_CUDA(h, cudaMemcpy(out, it->second.p, (size_t)len * 8, cudaMemcpyDeviceToHost));
  return ADAPROX_OK;
}

extern "C" int adaprox_vector_free(adaprox_handle h, adaprox_id vec) {
  if (!h) return ADAPROX_ERR_INVALID;
  auto it = h->vecs.find(vec);
  if (it == h->vecs.end()) return fail(h, ADAPROX_ERR_INVALID, "vector_free: unknown id");
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  cudaFree(it->second.p);
  h->vecs.erase(it);
  return ADAPROX_OK;
}

extern "C" int adaprox_matrix_set_shard(adaprox_handle h, adaprox_id mat, int64_t m_global, int64_t row0) {
  if (!h) return ADAPROX_ERR_INVALID;
  HostMatrix* hm;
  int rc = get_mat(h, mat, &hm);
  if (rc) return rc;
  if (m_global < hm->d.m || row0 < 0 || row0 + hm->d.m > m_global) return fail(h, ADAPROX_ERR_INVALID, "matrix_set_shard: bad row range");
  hm->m_global = m_global; hm->row0 = row0; hm->sharded = (m_global != hm->d.m);
  return ADAPROX_OK;
}

// ---------------------------------------------------------------------------
// operator protocol, one call each
// ---------------------------------------------------------------------------
extern "C" int adaprox_mul(adaprox_handle h, adaprox_id mat, const double* x, double* out) {
  if (!h || !x || !out) return fail(h, ADAPROX_ERR_INVALID, "mul: bad arguments");
  HostMatrix* hm;
  int rc = get_mat(h, mat, &hm);
  if (rc) return rc;
  AP_CUDA(h, cudaSetDevice(h->device));
  const DMat& M = hm->d;
  if ((rc = ws_reset(h, ws_size_doubles(M.n) + ws_size_doubles(M.m)))) return rc;
  double* dx = ws_doubles(h, M.n);
  double* dz = ws_doubles(h, M.m);
  AP_CUDA(h, cudaMemcpyAsync(dx, x, (size_t)M.n * 8, cudaMemcpyHostToDevice, h->stream));
  if ((rc = op_mul_dev(h, M, dx, dz))) return rc;
  AP_CUDA(h, cudaMemcpyAsync(out, dz, (size_t)M.m * 8, cudaMemcpyDeviceToHost, h->stream));
  AP_CUDA(h, cudaStreamSynchronize(h->stream));
  return ADAPROX_OK;
}

extern "C" int adaprox_amul(adaprox_handle h, adaprox_id mat, const double* y, double* out) {
  if (!h || !y || !out) return fail(h, ADAPROX_ERR_INVALID, "amul: bad arguments");
  HostMatrix* hm;
  int rc = get_mat(h, mat, &hm);
  if (rc) return rc;
  AP_CUDA(h, cudaSetDevice(h->device));
  const DMat& M = hm->d;
  if ((rc = ws_reset(h, ws_size_doubles(M.n) + ws_size_doubles(M.m)))) return rc;
  double* dy = ws_doubles(h, M.m);
  double* dg = ws_doubles(h, M.n);
  AP_CUDA(h, cudaMemcpyAsync(dy, y, (size_t)M.m * 8, cudaMemcpyHostToDevice, h->stream));
  if ((rc = op_amul_dev(h, M, dy, dg))) return rc;
  if (hm->sharded && h->comm) { if ((rc = comm_allreduce_sum(h, dg, M.n))) return rc; }
  AP_CUDA(h, cudaMemcpyAsync(out, dg, (size_t)M.n * 8, cudaMemcpyDeviceToHost, h->stream));
  AP_CUDA(h, cudaStreamSynchronize(h->stream));
  return ADAPROX_OK;
}

extern "C" int adaprox_eval_f(adaprox_handle h, const adaprox_problem* p, const double* x, double* f_x, double* grad) {
  if (!h || !p || !x) return fail(h, ADAPROX_ERR_INVALID, "eval_f: bad arguments");
  AP_CUDA(h, cudaSetDevice(h->device));
  DProblem P;
  HostMatrix *fm, *am;
  adaprox_problem q = *p;
  q.A_mat = 0;
  int rc = fill_problem(h, &q, &P, &fm, &am);
  if (rc) return rc;
  if (fm && fm->sharded) return fail(h, ADAPROX_ERR_UNSUPPORTED, "eval_f on a row shard: use the solver entry points");
  const int64_t mf = std::max<int64_t>(P.F.kind != MAT_NONE ? P.F.m : 0, P.n);
  const int64_t nfu = (P.f_kind == ADAPROX_F_QUADRATIC_GRAM) ? P.F.n : 1;
  if ((rc = ws_reset(h, 2 * ws_size_doubles(P.n) + ws_size_doubles(mf) + ws_size_doubles(nfu) + ws_size_doubles((int64_t)kMaxRed * h->grid) + ws_size_doubles(4)))) return rc;
  double* dx = ws_doubles(h, P.n);
  double* dg = ws_doubles(h, P.n);
  DWork W{};
  W.r = ws_doubles(h, mf);
  W.fu = ws_doubles(h, nfu);
  W.red = ws_doubles(h, (int64_t)kMaxRed * h->grid);
  double* scal = ws_doubles(h, 4);
  AP_CUDA(h, cudaMemcpyAsync(dx, x, (size_t)P.n * 8, cudaMemcpyHostToDevice, h->stream));
  OpArgs a{}; a.op = OP_EVALF; a.P = P; a.in = dx; a.out = dg; a.scal = scal; a.want_grad = grad ? 1 : 0;
  if ((rc = run_ops(h, a, W))) return rc;
  double fx = 0.0;
  AP_CUDA(h, cudaMemcpyAsync(&fx, scal, 8, cudaMemcpyDeviceToHost, h->stream));
  if (grad) AP_CUDA(h, cudaMemcpyAsync(grad, dg, (size_t)P.n * 8, cudaMemcpyDeviceToHost, h->stream));
  AP_CUDA(h, cudaStreamSynchronize(h->stream));
  if (f_x) *f_x = fx;
  return ADAPROX_OK;
}

extern "C" int adaprox_prox_eval(adaprox_handle h, const adaprox_prox* g, const double* x, int64_t len, double gamma,
                                 double* y, double* g_y) {
  if (!h || !g || !x || !y || len <= 0) return fail(h, ADAPROX_ERR_INVALID, "prox_eval: bad arguments");
  AP_CUDA(h, cudaSetDevice(h->device));
  DProx px;
  int rc = fill_prox(h, *g, len, &px);
  if (rc) return rc;
  if ((rc = ws_reset(h, 2 * ws_size_doubles(len) + ws_size_doubles((int64_t)kMaxRed * h->grid) + ws_size_doubles(4)))) return rc;
  double* dx = ws_doubles(h, len);
  double* dy = ws_doubles(h, len);
  DWork W{};
  W.red = ws_doubles(h, (int64_t)kMaxRed * h->grid);
  double* scal = ws_doubles(h, 4);
  AP_CUDA(h, cudaMemcpyAsync(dx, x, (size_t)len * 8, cudaMemcpyHostToDevice, h->stream));
  OpArgs a{}; a.op = OP_PROX; a.px = px; a.gamma = gamma; a.len = len; a.in = dx; a.out = dy; a.scal = scal;
  if ((rc = run_ops(h, a, W))) return rc;
  double gy = 0.0;
  AP_CUDA(h, cudaMemcpyAsync(&gy, scal, 8, cudaMemcpyDeviceToHost, h->stream));
  AP_CUDA(h, cudaMemcpyAsync(y, dy, (size_t)len * 8, cudaMemcpyDeviceToHost, h->stream));
  AP_CUDA(h, cudaStreamSynchronize(h->stream));
  if (g_y) *g_y = gy;
  return ADAPROX_OK;
}

static void fill_opts(const adaprox_options* o, DOpts* d) {
  d->solver = o->solver; d->rule = o->rule; d->gamma = o->gamma; d->t = o->t; d->norm_A = o->norm_A; d->delta = o->delta;
  d->Theta = o->Theta; d->xi = o->xi; d->nu = o->nu; d->r = o->r; d->R = o->R; d->eta = o->eta; d->shrink = o->shrink;
  d->sigma = o->sigma; d->muf = o->muf; d->mug = o->mug; d->theta = o->theta; d->gamma_max = o->gamma_max; d->phi = o->phi;
  d->tol = o->tol; d->maxit = o->maxit; d->max_records = o->max_records; d->want_objective = o->want_objective;
}

extern "C" int adaprox_stepsize(const adaprox_options* o, double gamma1, double gamma0_or_rho, double dgg, double dgx,
                                double dxx, double* gamma, double* sigma, double* state1) {
  if (!o || !gamma || !sigma) return ADAPROX_ERR_INVALID;
  DOpts d{};
  fill_opts(o, &d);
  double g = 0, s = 0, s0 = gamma1, s1 = gamma0_or_rho;
  rule_step(d, dgg, dgx, dxx, g, s, s0, s1);
  *gamma = g; *sigma = s;
  if (state1) *state1 = s1;
  return ADAPROX_OK;
}

// ---------------------------------------------------------------------------
// solvers
// ---------------------------------------------------------------------------
static int validate_options(adaprox_ctx* h, const adaprox_problem* p, const adaprox_options* o) {
  if (o->maxit < 0) return fail(h, ADAPROX_ERR_INVALID, "maxit must be >= 0");
  const bool pd = (o->solver == ADAPROX_S_ADAPTIVE_PRIMAL_DUAL || o->solver == ADAPROX_S_LINESEARCH_PRIMAL_DUAL ||
                   o->solver == ADAPROX_S_MALITSKY_POCK);
  if (pd && p->A_mat == 0) return fail(h, ADAPROX_ERR_INVALID, "primal-dual solvers need A (use adaptive_proxgrad for A = 0)");
  if (!pd && p->A_mat != 0) return fail(h, ADAPROX_ERR_INVALID, "proximal-gradient solvers take no A");
  // g = NormL2 or a conjugate g needs one reduction before the prox (src/AdaProx.jl:332,361 accept any prox-able g): the
  // adaptive primal-dual / proximal-gradient loops have it (solver_pd.cuh primal_step); the comparison baselines do not.
  const bool pd_loop = (o->solver == ADAPROX_S_ADAPTIVE_PRIMAL_DUAL || o->solver == ADAPROX_S_ADAPTIVE_PROXGRAD ||
                        o->solver == ADAPROX_S_LINESEARCH_PRIMAL_DUAL);
  if ((p->g.kind == ADAPROX_P_NORM_L2 || p->g.conjugate) && !pd_loop)
    return fail(h, ADAPROX_ERR_UNSUPPORTED, "g = NormL2 / conjugate g: supported by adaptive_primal_dual, adaptive_proxgrad, fixed_proxgrad, condat_vu and AdaPDM+ only");
  if (p->A_mat != 0 && p->h.conjugate && !pd_loop)
    return fail(h, ADAPROX_ERR_UNSUPPORTED, "h given as a conjugate: supported by adaptive_primal_dual, condat_vu and AdaPDM+ only");
  switch (o->solver) {
    case ADAPROX_S_ADAPTIVE_PRIMAL_DUAL:
    case ADAPROX_S_ADAPTIVE_PROXGRAD:
      if (o->rule < ADAPROX_RULE_FIXED || o->rule > ADAPROX_RULE_OUR_PLUS) return fail(h, ADAPROX_ERR_INVALID, "unknown stepsize rule");
      if (!(o->gamma > 0)) return fail(h, ADAPROX_ERR_INVALID, "you must provide gamma > 0 if norm_A = 0");   // :246,:288
      break;
    case ADAPROX_S_LINESEARCH_PRIMAL_DUAL:
      if (!(o->eta > 0)) return fail(h, ADAPROX_ERR_INVALID, "eta must be positive");                          // :481
      if (!(o->Theta > o->delta + 1)) return fail(h, ADAPROX_ERR_INVALID, "must be Theta > (delta + 1)");       // :482
      if (!(o->gamma <= 1.0 / (2.0 * o->Theta * o->t * o->eta))) return fail(h, ADAPROX_ERR_INVALID, "gamma is too large");   // :488
      break;
    case ADAPROX_S_BACKTRACKING_PROXGRAD:
    case ADAPROX_S_BACKTRACKING_NESTEROV:
      if (!(o->gamma > 0) || !(o->shrink > 0 && o->shrink < 1)) return fail(h, ADAPROX_ERR_INVALID, "backtracking needs gamma0 > 0, 0 < shrink < 1");
      break;
    case ADAPROX_S_FIXED_NESTEROV: {
      if (!(o->gamma > 0)) return fail(h, ADAPROX_ERR_INVALID, "fixed_nesterov needs gamma (or Lf) > 0");
      const double mu = o->muf + o->mug, q = o->gamma * mu / (1 + o->gamma * o->mug);
      if (!(q < 1)) return fail(h, ADAPROX_ERR_INVALID, "fixed_nesterov: q < 1 violated");                    // :110
    } break;
    case ADAPROX_S_AGRAAL:          // gamma <= 0 means `gamma0 = nothing` (:168-170)
      if (!(o->phi > 1)) return fail(h, ADAPROX_ERR_INVALID, "agraal needs phi > 1");
      break;
    case ADAPROX_S_MALITSKY_POCK:
      if (!(o->sigma > 0)) return fail(h, ADAPROX_ERR_INVALID, "malitsky_pock needs sigma > 0");
      break;
    default:
      return fail(h, ADAPROX_ERR_INVALID, "unknown solver");
  }
  return ADAPROX_OK;
}

// Single-pass fused AdaPGM (solver_fused.cuh): cluster launch sized to full residency.  Returns 1 if the configuration is not
// eligible (the caller falls through to the two-pass kernel), < 0 on error.
constexpr int64_t kSmallProblemBytes = 8 << 20;   // fewer matrix bytes than this: latency-bound, one CTA per SM (see solver_grid)
static int fused_cluster_size(const DProblem& P) { return (int)((P.F.ld + kFCols - 1) / kFCols); }
// `rows` = the row count the decision is based on: the matrix's own rows on one GPU; for a row shard the GLOBAL rows divided
// by the number of ranks, so that every rank takes the same decision even when the shards differ by a row (ranks that
// disagreed would wait for each other in different collectives).  ADAPROX_FUSED is read per process: set it on all ranks.
static bool fused_eligible(const adaprox_options* o, const DProblem& P, int64_t rows) {
  const char* e = std::getenv("ADAPROX_FUSED");
  if (e && std::strcmp(e, "0") == 0) return false;
  if (o->solver != ADAPROX_S_ADAPTIVE_PROXGRAD || P.f_kind != ADAPROX_F_LEAST_SQUARES || P.F.kind != MAT_DENSE) return false;
  if (P.g.kind == ADAPROX_P_NORM_L2 || P.g.conjugate) return false;       // needs a reduction before the prox: general kernel
  const bool force = e && std::strcmp(e, "1") == 0;
  // Where the single sweep pays (tools/l2_fit_ab.py, us per iteration two-pass / fused, round 2): rows of at least half a 64 KB
  // stage -- 1000 x 4096: 43.8 / 34.7, 4000 x 4096: 72.3 / 59.0, 300 x 16384: 55.9 / 32.1, 200 x 65536: 60.7 / 39.5,
  // 8000 x 8192: 185 / 125 -- and not with short rows, where a CTA's stage holds a fraction of a row: 4000 x 1000: 42.8 / 55.7,
  // 20000 x 1000: 95.4 / 157, 8000 x 2048: 66.7 / 97.2 (1000 x 2048: 36.1 / 34.1, a tie).
  if (!force && (P.F.ld < kFCols / 2 || rows * P.F.ld * 8 < kSmallProblemBytes)) return false;
  return fused_cluster_size(P) <= kFMaxCluster;
}
static int fused_config(adaprox_ctx* h, const void* kernel, int C, bool cooperative, cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attrs, int* Q) {
  AP_CUDA(h, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFRingBytes));
  AP_CUDA(h, cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  *cfg = cudaLaunchConfig_t{};
  cfg->gridDim = dim3(C, 1, 1);
  cfg->blockDim = dim3(kFThreads, 1, 1);
  cfg->dynamicSmemBytes = kFRingBytes;
  cfg->stream = h->stream;
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = C; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
  cfg->attrs = attrs;
  cfg->numAttrs = 1;
  int q = 0;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&q, kernel, cfg);
  if (e != cudaSuccess || q < 1) { cudaGetLastError(); return 1; }
  *Q = q;
  cfg->gridDim = dim3(C * q, 1, 1);
  // The kernel synchronises the grid with its own arrival-counter barrier (GridBar), so every CTA must be resident:
  // the persistent single-GPU solve adds the cooperative attribute, which makes the runtime refuse the launch
  // (cudaErrorCooperativeLaunchTooLarge) instead of letting it spin.  ADAPROX_FUSED_NONCOOP=1: plain cluster launch
  // (Nsight Compute cannot replay cooperative + cluster); the barrier's spin is bounded either way.
  if (cooperative && !std::getenv("ADAPROX_FUSED_NONCOOP")) {
    attrs[1].id = cudaLaunchAttributeCooperative;
    attrs[1].val.cooperative = 1;
    cfg->numAttrs = 2;
  }
  return ADAPROX_OK;
}

// Plan of one fused solve: launch configuration, chunk size of the dynamic schedule, workspace.
struct FusedPlan {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attrs[2];
  int Q = 0;                 // resident clusters
  int G = 0;                 // CTAs = C * Q
  FusedArgs fa{};
};
// returns 0 = ok, 1 = not possible on this device (fall back to the two-pass kernels), < 0 = error
static int fused_plan(adaprox_ctx* h, const void* kernel, const DProblem& P, bool cooperative, FusedPlan* pl) {
  pl->fa = FusedArgs{};
  pl->fa.C = fused_cluster_size(P);
  int rc = fused_config(h, kernel, pl->fa.C, cooperative, &pl->cfg, pl->attrs, &pl->Q);
  if (rc) return rc;
  pl->G = pl->fa.C * pl->Q;
  // persistent launches (`cooperative`): tagged chunk dispenser, and helper CTAs on the SMs the clusters leave idle -- virtual
  // clusters of 16 (solver_fused_helper.cuh).  ADAPROX_HELPERS=V overrides (0: none), ADAPROX_HELPER_ROWS=R rows per helper batch.
  pl->fa.tagged = cooperative ? 1 : 0;
  pl->fa.hV = 0; pl->fa.hR = 16;
#if ADAPROX_FUSED_VARIANT == 1
  if (cooperative && pl->fa.C == kFMaxCluster) {
    // Default: as many virtual clusters as fit the SMs the clusters leave idle (2 on B200).  Measured (same box, profiles/r02_notes.md
    // section 7): 87.2-88.4 vs 84.1-85.5 it/s -- the helpers take ~20 % of the chunks, but the sweep kernel runs into the 1000 W
    // power cap, so the clusters slow down by most of that.  ADAPROX_HELPERS=0 disables them.
    int V = (h->sm_count - pl->G) / kFMaxCluster;
    // ... for LONG sweeps only: with 8192 rows per GPU (the N = 8 shard) they cost 2 % (686.7 -> 673.4 it/s on one GPU with m = 8192: a helper
    // chunk takes 1.7x a cluster's, the sweep is 1.4 ms), with 16384 rows they are neutral, with 65536 rows they give +5.5 % (84.3 -> 88.9)
    if (P.F.m / std::max(pl->Q, 1) < 4096) V = 0;
    if (const char* e = std::getenv("ADAPROX_HELPERS")) V = std::min((h->sm_count - pl->G) / kFMaxCluster, std::max(0, std::atoi(e)));
    pl->fa.hV = h->resident_dev ? std::min(V, 4) : 0;
    pl->fa.hHold = (int)(1.7 * pl->Q) + 1;
    if (const char* e = std::getenv("ADAPROX_HELPER_HOLD")) pl->fa.hHold = std::max(0, std::atoi(e));
    pl->fa.resident = h->resident_dev;
    pl->fa.resident_seq = ++h->solve_seq;
    if (const char* e = std::getenv("ADAPROX_HELPER_ROWS")) pl->fa.hR = std::min(64, std::max(1, std::atoi(e)));
  }
#endif
  pl->fa.npadf = (int64_t)pl->fa.C * kFCols;
  // chunk size: k chunks per cluster, k = 8 for long sweeps and 4 for short ones (row shards at N = 8).  A chunk boundary costs a
  // cluster barrier, an atomic, the reload of x and a refill of the 3-slot ring (~10 us); chunks of ceil(m / (Q k)) rows make the
  // number of chunks a multiple of the cluster count, so equally fast clusters finish together, while a cluster that is 2x slower
  // than the rest -- seen on some parts -- only delays the sweep by 12 % (k = 8) / 25 % (k = 4) instead of 100 %.
  const int kper = (P.F.m / std::max(pl->Q, 1) >= 4096) ? 8 : 4;
  int64_t pw = (P.F.m + (int64_t)pl->Q * kper - 1) / ((int64_t)pl->Q * kper);
  pw = pw < 16 ? 16 : pw;
  if (const char* e = std::getenv("ADAPROX_FUSED_CHUNK")) { const int v = std::atoi(e); if (v >= 1) pw = v; }
  pl->fa.chunk_rows = (int)pw;
  pl->fa.nchunks = (int)((P.F.m + pw - 1) / pw);
  return 0;
}
static size_t fused_ws_bytes(const FusedPlan& pl) {
  return ws_size_doubles((int64_t)pl.fa.nchunks * pl.fa.npadf) + ws_size_doubles(pl.fa.nchunks) + 5 * ws_size_doubles(1) + ws_size_doubles(12) +
         ws_size_doubles((int64_t)4 * 2 * 64 * kFMaxCluster * kFGWarps) +
         ws_size_doubles(4 * 256) + ws_size_doubles(5 * kFTraceRows);
}
static int fused_ws_alloc(adaprox_ctx* h, FusedPlan* pl) {
  FusedArgs& fa = pl->fa;
  fa.gpartf = ws_doubles(h, (int64_t)fa.nchunks * fa.npadf);
  fa.fpart = ws_doubles(h, fa.nchunks);
  fa.bar = reinterpret_cast<unsigned long long*>(ws_doubles(h, 1));
  fa.next = reinterpret_cast<unsigned long long*>(ws_doubles(h, 1));
  fa.err = reinterpret_cast<int*>(ws_doubles(h, 1));
  fa.go = reinterpret_cast<unsigned long long*>(ws_doubles(h, 1));
  fa.done = reinterpret_cast<unsigned long long*>(ws_doubles(h, 1));
  fa.hsync = reinterpret_cast<unsigned long long*>(ws_doubles(h, 12));
  fa.hxch = ws_doubles(h, (int64_t)4 * 2 * 64 * kFMaxCluster * kFGWarps);
  AP_CUDA(h, cudaMemsetAsync(fa.go, 0, 8, h->stream));
  AP_CUDA(h, cudaMemsetAsync(fa.done, 0, 8, h->stream));
  AP_CUDA(h, cudaMemsetAsync(fa.hsync, 0, 12 * 8, h->stream));
  AP_CUDA(h, cudaMemsetAsync(fa.bar, 0, 8, h->stream));
  AP_CUDA(h, cudaMemsetAsync(fa.next, 0, 8, h->stream));
  AP_CUDA(h, cudaMemsetAsync(fa.err, 0, 8, h->stream));
  unsigned long long* lat = reinterpret_cast<unsigned long long*>(ws_doubles(h, 4 * 256));
  if (std::getenv("ADAPROX_FUSED_LAT")) {
    fa.lat = lat;
    AP_CUDA(h, cudaMemsetAsync(fa.lat, 0, 4 * 256 * 8, h->stream));
  }
  unsigned long long* trace = reinterpret_cast<unsigned long long*>(ws_doubles(h, 5 * kFTraceRows));
#ifdef ADAPROX_FUSED_TRACE
  fa.trace = trace;
  AP_CUDA(h, cudaMemsetAsync(fa.trace, 0, 5 * kFTraceRows * 8, h->stream));
#else
  (void)trace;
#endif
  return ADAPROX_OK;
}
// did a grid barrier of the fused kernel time out?  (GridBar; results are garbage then)
static int fused_check(adaprox_ctx* h, const FusedPlan& pl) {
  int e = 0;
  AP_CUDA(h, cudaMemcpy(&e, pl.fa.err, sizeof(int), cudaMemcpyDeviceToHost));
  if (e) return fail(h, ADAPROX_ERR_CUDA, "fused sweep kernel: grid barrier timed out (its CTAs were not all resident -- another kernel on this GPU?)");
  return ADAPROX_OK;
}
static void fused_print_probe(const FusedPlan& pl, const char* what) {
#ifdef ADAPROX_FUSED_TRACE
  if (pl.fa.trace) {        // pipeline timeline of CTA 0, second chunk of the last sweep (SM clock cycles relative to the first issue)
    std::vector<unsigned long long> T((size_t)5 * kFTraceRows);
    cudaMemcpy(T.data(), pl.fa.trace, T.size() * 8, cudaMemcpyDeviceToHost);
    const unsigned long long t0 = T[0];
    std::fprintf(stderr, "[adaprox %s trace: row issue full dot_done exch_done upd_done (cycles)]\n", what);
    for (int i = 0; i < kFTraceRows; ++i) {
      if (!T[(size_t)1 * kFTraceRows + i]) break;
      std::fprintf(stderr, "T %d %lld %lld %lld %lld %lld\n", i, (long long)(T[i] - t0), (long long)(T[kFTraceRows + i] - t0),
                   (long long)(T[2 * kFTraceRows + i] - t0), (long long)(T[3 * kFTraceRows + i] - t0), (long long)(T[4 * kFTraceRows + i] - t0));
    }
  }
#endif
  if (!pl.fa.lat) return;
  unsigned long long L4[4 * 256];
  cudaMemcpy(L4, pl.fa.lat, sizeof(L4), cudaMemcpyDeviceToHost);
  std::fprintf(stderr, "[adaprox %s: chunk_rows=%d nchunks=%d; per cluster: chunks taken (all sweeps) / last sweep us @ first smid]", what,
               pl.fa.chunk_rows, pl.fa.nchunks);
  for (int i = 0; i < pl.G; i += pl.fa.C)
    std::fprintf(stderr, " %llu/%.0f@%llu", L4[4 * i], (double)L4[4 * i + 1] * 1e-3, L4[4 * i + 3]);
  std::fprintf(stderr, "\n");
}

// Small dense least squares: the matrix fits the distributed shared memory of one 16-CTA cluster (solver_resident.cuh).
// ADAPROX_RESIDENT=0 disables it (A/B against the persistent grid kernel).
static bool resident_eligible(adaprox_ctx* h, const adaprox_options* o, const DProblem& P, ResidentArgs* ra, size_t* smem) {
  const char* e = std::getenv("ADAPROX_RESIDENT");
  if (e && std::strcmp(e, "0") == 0) return false;
  const char* ef = std::getenv("ADAPROX_FUSED");
  if (ef && std::strcmp(ef, "1") == 0) return false;       // the sweep kernel was requested explicitly
  if (o->solver != ADAPROX_S_ADAPTIVE_PROXGRAD || P.f_kind != ADAPROX_F_LEAST_SQUARES || P.F.kind != MAT_DENSE) return false;
  if (P.g.kind == ADAPROX_P_NORM_L2 || P.g.conjugate) return false;
  if (P.F.ld > kRMaxLd || P.n < 1) return false;
  ra->rows_cap = (int)((P.F.m + kRCluster - 1) / kRCluster);
  ra->slice = (int)((P.n + kRCluster - 1) / kRCluster);
  *smem = resident_smem_bytes(ra->rows_cap, P.F.ld);
  int max_optin = 0;
  if (cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device) != cudaSuccess) return false;
  return *smem + 1024 <= (size_t)max_optin;
}

// Dense least squares with short rows that fits the shared memory of all SMs together (solver_gridres.cuh): the reference's
// 500 x 1000 and 4000 x 1000 lasso runs.  ADAPROX_GRIDRES=0 disables it (A/B against the persistent grid kernel).
static const void* gridres_kernel(int solver) {
  switch (solver) {
    case ADAPROX_S_ADAPTIVE_PROXGRAD: return (const void*)k_adapgm_gridres<0>;
    case ADAPROX_S_FIXED_NESTEROV: return (const void*)k_adapgm_gridres<1>;
    case ADAPROX_S_AGRAAL: return (const void*)k_adapgm_gridres<2>;
    case ADAPROX_S_BACKTRACKING_PROXGRAD: return (const void*)k_adapgm_gridres<3>;
    case ADAPROX_S_BACKTRACKING_NESTEROV: return (const void*)k_adapgm_gridres<4>;
    default: return nullptr;
  }
}
static bool gridres_eligible(adaprox_ctx* h, const adaprox_options* o, const DProblem& P, GridResArgs* ga, size_t* smem) {
  const char* e = std::getenv("ADAPROX_GRIDRES");
  if (e && std::strcmp(e, "0") == 0) return false;
  const char* ef = std::getenv("ADAPROX_FUSED");
  if (ef && std::strcmp(ef, "1") == 0) return false;       // the sweep kernel was requested explicitly
  const void* kernel = gridres_kernel(o->solver);          // AdaPGM / fixed-step PGM, fixed_nesterov, agraal, the two backtracking methods
  if (!kernel || P.f_kind != ADAPROX_F_LEAST_SQUARES || P.F.kind != MAT_DENSE || P.A.kind != MAT_NONE) return false;
  if (P.g.kind == ADAPROX_P_NORM_L2 || P.g.conjugate) return false;
  if (P.F.ld > kGMaxLd || P.n < 1 || P.F.m < 1) return false;
  const int G = h->sm_count;
  if (G < 1 || G > h->grid || G > 32 * kGMaxP) return false;   // the matrix's partial buffer has h->grid rows
  ga->rows_cap = (int)((P.F.m + G - 1) / G);
  ga->slice = (int)((P.n + G - 1) / G);
  ga->row_ctas = (int)((P.F.m + ga->rows_cap - 1) / ga->rows_cap);
  if (ga->slice > kGWarps) return false;
  int max_optin = 0;
  if (cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device) != cudaSuccess) return false;
  // the CTA's copy of x sits in shared memory when the rows leave room for it (4000 x 1000: 28 rows of 8064 B leave 6.6 KB), else in global memory
  ga->x_in_smem = gridres_smem_bytes(ga->rows_cap, P.F.ld, true) + 1024 <= (size_t)max_optin ? 1 : 0;
  *smem = gridres_smem_bytes(ga->rows_cap, P.F.ld, ga->x_in_smem != 0);
  if (*smem + 1024 > (size_t)max_optin) return false;
  int per_sm = 0;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kGThreads, *smem) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    return false;
  }
  return true;
}

// CTAs of a persistent solver launch.  Every phase boundary is a grid barrier whose cost grows with the number of CTAs, and
// every scalar is rebuilt from one partial per CTA: a problem whose matrices are small enough to be latency-bound (a few
// microseconds per phase) runs faster on fewer CTAs; a streaming problem wants all of them (2 per SM keep ~96 KB of bulk
// copies in flight per SM).  The matrix partial buffers are sized for h->grid, so any smaller grid is valid.
// Row-sharded solves always use h->grid: all ranks must reduce the replicated vectors in the same order.
static int64_t mat_bytes(const DMat& M) {
  if (M.kind == MAT_DENSE) return M.m * M.ld * 8;
  if (M.kind == MAT_CSR) return 2 * M.nnz * 12;
  return 0;
}
static int solver_grid(adaprox_ctx* h, const DProblem& P) {
  if (const char* e = std::getenv("ADAPROX_GRID")) {
    const int v = std::atoi(e);
    if (v >= 1) return std::min(v, h->grid);
  }
  const int64_t bytes = mat_bytes(P.F) + mat_bytes(P.A);
  // measured (tools/grid_sweep.py, profiles/r01_grid_sweep.jsonl): 400x1000 lasso (3.2 MB) 38.0 us/iteration on 296 CTAs,
  // 30.8 on 148, 31.9 on 74, 40.5 on 37; already at 36 MB (rcv1-shaped CSR) and 64 MB (2048x4096 dense) 296 CTAs win.
  // The primal-dual loops keep the full grid: their row phases and the one-row tiles of a narrow A want more CTAs
  // (cpusmall-shaped LAD 8192x13, AdaPDM+: 74.9 us on 296 CTAs, 83.7 on 148; mushrooms-shaped CSR logreg, AdaPGM: 36.9 vs 33.9).
  if (P.A.kind == MAT_NONE && bytes < kSmallProblemBytes) return std::min(h->grid, h->sm_count);
  return h->grid;
}

extern "C" int adaprox_solve(adaprox_handle h, const adaprox_problem* p, const adaprox_options* o, const double* x0,
                             const double* y0, double* x_out, double* y_out, adaprox_record* records, adaprox_result* res) {
  if (!h || !p || !o || !x0 || !x_out || !res) return fail(h, ADAPROX_ERR_INVALID, "solve: bad arguments");
  AP_CUDA(h, cudaSetDevice(h->device));
  int rc;
  if ((rc = validate_options(h, p, o))) return rc;
  DProblem P;
  HostMatrix *fm, *am;
  if ((rc = fill_problem(h, p, &P, &fm, &am))) return rc;
  DOpts O{};
  fill_opts(o, &O);
  if (!records) O.max_records = 0;
  const bool sharded = (fm && fm->sharded) || (am && am->sharded);
  // Row-sharded linear map A of the primal-dual loops (AdaPDM / Condat-Vu / AdaPDM+; f without a sharded matrix): the persistent
  // kernel itself all-reduces A'y and the dual sums over NVLink peer memory -- needs the exchange blocks (adaprox_p2p_*).
  // ... and a row-sharded Q of a Quadratic smooth term (dual SVM): gradient rows and value sums are gathered the same way.
  const bool pd_solver = o->solver == ADAPROX_S_ADAPTIVE_PRIMAL_DUAL || o->solver == ADAPROX_S_LINESEARCH_PRIMAL_DUAL;
  const bool f_quad_shard = fm && fm->sharded && (P.f_kind == ADAPROX_F_QUADRATIC || P.f_kind == ADAPROX_F_QUADRATIC_GRAM);
  const bool f_ok = !(fm && fm->sharded) || f_quad_shard;
  // ... and the backtracking / Nesterov / aGRAAL baselines on a row-sharded least-squares or Quadratic term.
  const bool pg_family = o->solver == ADAPROX_S_BACKTRACKING_PROXGRAD || o->solver == ADAPROX_S_BACKTRACKING_NESTEROV ||
                         o->solver == ADAPROX_S_FIXED_NESTEROV || o->solver == ADAPROX_S_AGRAAL;
  const bool f_ls_shard = fm && fm->sharded && P.f_kind == ADAPROX_F_LEAST_SQUARES && P.F.kind == MAT_DENSE;
  const bool sharded_pd = sharded && ((f_ok && pd_solver && ((am && am->sharded) || f_quad_shard)) ||
                                      (f_ok && o->solver == ADAPROX_S_ADAPTIVE_PROXGRAD && f_quad_shard) ||
                                      (pg_family && (f_ls_shard || f_quad_shard) && !(am && am->sharded)));
  // Row-sharded AdaPGM / fixed-step PGM on a dense least-squares shard with the exchange blocks attached: the single-sweep kernel
  // stays resident for the whole solve and all-reduces inside its loop (solver_fused.cuh, `sharded`): one launch per rank.
  // ADAPROX_SHARDED_LAUNCHES=1 keeps the round-1 structure (sweep-only launch + two small kernels per iteration) for A/B; without
  // the exchange blocks (or under ADAPROX_NO_P2P) the split-phase path with ncclAllReduce runs.  All of these are read per process
  // and must agree on every rank.
  const bool sharded_fused = sharded && !sharded_pd && f_ls_shard && o->solver == ADAPROX_S_ADAPTIVE_PROXGRAD && h->comm &&
                             p2p_ready(h, P.n + 2) && !std::getenv("ADAPROX_NO_P2P") && !std::getenv("ADAPROX_SHARDED_LAUNCHES") &&
                             fused_eligible(o, P, (fm->m_global + comm_nranks(h) - 1) / comm_nranks(h));
  if (sharded_pd) {
    if (!p2p_ready(h, std::max<int64_t>(std::max<int64_t>(P.n, P.f_kind == ADAPROX_F_QUADRATIC_GRAM ? P.F.n : 0), 8)))
      return fail(h, ADAPROX_ERR_COMM, "row-sharded primal-dual solve: attach the peer exchange blocks first (adaprox_p2p_export / adaprox_p2p_attach)");
    p2p_fill(h, &P.p2p);
    P.A_sharded = (am && am->sharded) ? 1 : 0;
    P.F_sharded = (f_quad_shard || (pg_family && f_ls_shard)) ? 1 : 0;
  } else if (sharded && !sharded_fused) {
    return solve_sharded(h, p, o, P, O, fm, am, x0, y0, x_out, y_out, records, res);
  }

  const int64_t n = P.n, md = std::max<int64_t>(P.md, 1);
  const int64_t mf = std::max<int64_t>(P.F.kind != MAT_NONE ? P.F.m : 0, n);
  const int64_t nrec = std::min<int64_t>(O.max_records, O.maxit);
  int G = sharded_pd ? h->grid : solver_grid(h, P);
  const int Gcoop = G;
  ResidentArgs rarg{};
  size_t rsmem = 0;
  const bool resident = !sharded && resident_eligible(h, o, P, &rarg, &rsmem);
  GridResArgs garg{};
  size_t gsmem = 0;
  bool gridres = !sharded && !resident && gridres_eligible(h, o, P, &garg, &gsmem);
  bool fused = sharded_fused || (!resident && !gridres && !sharded && fused_eligible(o, P, P.F.m));
  if (gridres) G = std::max(G, h->sm_count);
  FusedPlan fpl;
  if (fused) {
    rc = fused_plan(h, (const void*)k_adapgm_fused, P, true, &fpl);
    if (rc < 0) return rc;
    if (rc == 1 && sharded_fused) return fail(h, ADAPROX_ERR_CUDA, "row-sharded fused solve: no resident cluster configuration on this device");
    if (rc == 1) fused = false; else G = fpl.G;
  }
  FusedArgs& fa = fpl.fa;
  const int fQ = fpl.Q;
  const int64_t nfu = (P.f_kind == ADAPROX_F_QUADRATIC_GRAM) ? P.F.n : 1;
  const int64_t gr_x = gridres ? (garg.x_in_smem ? 1 : (int64_t)h->sm_count * P.F.ld) : 1;
  size_t need = (fused ? fused_ws_bytes(fpl) : 0) + (sharded_fused ? ws_size_doubles(n + 2) : 0) +
                (gridres ? ws_size_doubles(P.F.ld) + ws_size_doubles(h->sm_count) + ws_size_doubles(2 * h->sm_count) + ws_size_doubles(gr_x) : 0) +
                9 * ws_size_doubles(n) + 2 * ws_size_doubles(n + 8) + 6 * ws_size_doubles(md) + ws_size_doubles(mf) + ws_size_doubles(nfu) +
                ws_size_doubles((int64_t)kMaxRed * G) + ws_size_doubles((nrec * (int64_t)sizeof(adaprox_record) + 7) / 8) +
                ws_size_doubles((sizeof(DResult) + 7) / 8);
  if ((rc = ws_reset(h, need))) return rc;
  DWork W{};
  for (int k = 0; k < 3; ++k) W.xb[k] = ws_doubles(h, n);
  for (int k = 0; k < 2; ++k) W.gb[k] = ws_doubles(h, n);
  W.v = ws_doubles(h, n);
  for (int k = 0; k < 2; ++k) W.Aty[k] = ws_doubles(h, n + 8);          // + the row sums that ride along in the sharded exchange
  for (int k = 0; k < 3; ++k) W.aux[k] = ws_doubles(h, n);
  for (int k = 0; k < 2; ++k) W.yb[k] = ws_doubles(h, md);
  W.w = ws_doubles(h, md);
  for (int k = 0; k < 2; ++k) W.Axb[k] = ws_doubles(h, md);
  W.yout = ws_doubles(h, md);
  W.r = ws_doubles(h, mf);
  W.fu = ws_doubles(h, nfu);
  W.red = ws_doubles(h, (int64_t)kMaxRed * G);
  W.rec = nrec > 0 ? reinterpret_cast<adaprox_record*>(ws_doubles(h, (nrec * (int64_t)sizeof(adaprox_record) + 7) / 8)) : nullptr;
  W.res = reinterpret_cast<DResult*>(ws_doubles(h, (sizeof(DResult) + 7) / 8));
  W.xout = W.aux[2];
  O.max_records = nrec;
  if (fused && (rc = fused_ws_alloc(h, &fpl))) return rc;
  if (gridres) {
    garg.gfull = ws_doubles(h, P.F.ld);
    garg.fpart = ws_doubles(h, h->sm_count);
    garg.fpart2 = ws_doubles(h, 2 * h->sm_count);
    garg.xpriv = ws_doubles(h, gr_x);
  }
  if (sharded_fused) {
    fa.sh_gbuf = ws_doubles(h, n + 2);
    p2p_fill(h, &fa.p2p);
  }
  const bool phase_timing = std::getenv("ADAPROX_PHASE_TIMING") != nullptr;
  unsigned long long* d_ts = nullptr;
  struct DevFree { unsigned long long*& p; ~DevFree() { if (p) cudaFree(p); } } d_ts_guard{d_ts};   // released on every return path
  const int ts_iters = (int)std::min<int64_t>(O.maxit, 64);
  if (phase_timing && ts_iters > 0) {
    AP_CUDA(h, cudaMalloc(&d_ts, (size_t)ts_iters * 8 * sizeof(unsigned long long)));
    AP_CUDA(h, cudaMemsetAsync(d_ts, 0, (size_t)ts_iters * 8 * sizeof(unsigned long long), h->stream));
    W.tstamp = d_ts; W.tstamp_iters = ts_iters;
  }

  AP_CUDA(h, cudaMemcpyAsync(W.xb[0], x0, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
  if (P.md > 0) {
    if (y0) AP_CUDA(h, cudaMemcpyAsync(W.yb[0], y0, (size_t)P.md * 8, cudaMemcpyHostToDevice, h->stream));
    else AP_CUDA(h, cudaMemsetAsync(W.yb[0], 0, (size_t)P.md * 8, h->stream));
  }
  if (o->solver == ADAPROX_S_AGRAAL) {      // y0 carries agraal's second start point x0 (:154,165)
    if (!y0) return fail(h, ADAPROX_ERR_INVALID, "agraal needs the second start point (pass it as y0)");
    AP_CUDA(h, cudaMemcpyAsync(W.aux[0], y0, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
  }
  AP_CUDA(h, cudaMemsetAsync(W.red, 0, (size_t)kMaxRed * G * 8, h->stream));
  const int64_t launches0 = h->launches;
  bool helper_launched = false;
  AP_CUDA(h, cudaEventRecord(h->ev_h, h->stream));          // workspace initialised (what the helper launch waits for)
  AP_CUDA(h, cudaEventRecord(h->ev0, h->stream));
  void* args[] = {&P, &O, &W};
  void* fargs[] = {&P, &O, &W, &fa};
  switch (o->solver) {
    case ADAPROX_S_ADAPTIVE_PRIMAL_DUAL:
    case ADAPROX_S_ADAPTIVE_PROXGRAD:
      if (resident) {
        AP_CUDA(h, cudaFuncSetAttribute((const void*)k_adapgm_resident, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem));
        AP_CUDA(h, cudaFuncSetAttribute((const void*)k_adapgm_resident, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cudaLaunchConfig_t cfg{};
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = kRCluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.gridDim = dim3(kRCluster, 1, 1); cfg.blockDim = dim3(kRThreads, 1, 1); cfg.dynamicSmemBytes = rsmem; cfg.stream = h->stream;
        cfg.attrs = at; cfg.numAttrs = 1;
        void* rargs[] = {&P, &O, &W, &rarg};
        cudaError_t e = cudaLaunchKernelExC(&cfg, (const void*)k_adapgm_resident, rargs);
        if (e != cudaSuccess) return fail(h, ADAPROX_ERR_CUDA, std::string("resident cluster launch: ") + cudaGetErrorString(e));
        h->launches++;
        rc = ADAPROX_OK;
      } else if (gridres) {
        void* gargs[] = {&P, &O, &W, &garg};
        cudaError_t e = cudaLaunchCooperativeKernel(gridres_kernel(o->solver), dim3(h->sm_count), dim3(kGThreads), gargs, gsmem, h->stream);
        if (e == cudaSuccess) {
          h->launches++;
          rc = ADAPROX_OK;
        } else {
          // one CTA per SM with all of its shared memory: refused (cudaErrorCooperativeLaunchTooLarge) when something else holds SMs.
          // The persistent grid kernel needs far less per SM; the workspace above covers it.
          cudaGetLastError();
          gridres = false;
          rc = coop_launch(h, k_primal_dual<false>, args, Gcoop);
        }
      } else if (fused) {
        cudaError_t e = cudaLaunchKernelExC(&fpl.cfg, (const void*)k_adapgm_fused, fargs);
        if (e != cudaSuccess) return fail(h, ADAPROX_ERR_CUDA, std::string("fused cluster launch: ") + cudaGetErrorString(e));
        h->launches++;
#if ADAPROX_FUSED_VARIANT == 1
        if (fa.hV > 0) {
          // helper CTAs beside the clusters: a plain launch on a second stream that only waits for the workspace initialisation (ev_h,
          // recorded before the main launch); they land on the SMs the resident cluster kernel leaves idle.  Optional by construction
          // (roll call in the kernel), so a failure to launch them is not an error.
          // Gate: the helpers must not become resident before the cluster kernel is (its 16-CTA clusters need whole GPC slots; helper CTAs
          // scattered over the GPCs first would keep them from being placed -- seen as a 60 s stall).  The cluster kernel reports through
          // mapped host memory after its first grid barrier; wait for that (microseconds), at most 2 s.
          bool up = false;
          for (long spin = 0; spin < 20000000L; ++spin) {
            if (*(volatile unsigned long long*)h->resident_host == fa.resident_seq) { up = true; break; }
            if ((spin & 1023) == 1023 && cudaStreamQuery(h->stream) != cudaErrorNotReady) break;      // the solve is already over (or failed)
          }
          if (up && cudaFuncSetAttribute((const void*)k_adapgm_helper, cudaFuncAttributeMaxDynamicSharedMemorySize, kFRingBytes) == cudaSuccess &&
              cudaStreamWaitEvent(h->stream2, h->ev_h, 0) == cudaSuccess) {
            k_adapgm_helper<<<fa.hV * kFMaxCluster, kFThreads, kFRingBytes, h->stream2>>>(P, W, fa);
            if (cudaGetLastError() == cudaSuccess) { h->launches++; helper_launched = true; }
          } else {
            cudaGetLastError();
          }
        }
#endif
        rc = ADAPROX_OK;
      } else {
        rc = coop_launch(h, k_primal_dual<false>, args, Gcoop);
      }
      break;
    case ADAPROX_S_LINESEARCH_PRIMAL_DUAL:
      rc = coop_launch(h, k_primal_dual<true>, args, Gcoop);
      break;
    case ADAPROX_S_BACKTRACKING_PROXGRAD:
    case ADAPROX_S_BACKTRACKING_NESTEROV:
    case ADAPROX_S_FIXED_NESTEROV:
    case ADAPROX_S_AGRAAL:
      rc = 1;
      if (gridres) {            // the comparison methods on a small dense least-squares term: the grid-resident form (solver_gridres.cuh)
        void* gargs[] = {&P, &O, &W, &garg};
        if (cudaLaunchCooperativeKernel(gridres_kernel(o->solver), dim3(h->sm_count), dim3(kGThreads), gargs, gsmem, h->stream) == cudaSuccess) {
          h->launches++;
          rc = ADAPROX_OK;
        } else {
          cudaGetLastError();
          gridres = false;
        }
      }
      if (rc == 1) rc = coop_launch(h, k_proxgrad_family, args, Gcoop);
      break;
    case ADAPROX_S_MALITSKY_POCK:
      rc = coop_launch(h, k_malitsky_pock, args, Gcoop);
      break;
    default:
      return fail(h, ADAPROX_ERR_UNSUPPORTED, "solver has no device kernel yet");
  }
  if (rc) return rc;
  AP_CUDA(h, cudaEventRecord(h->ev1, h->stream));
  DResult dr{};
  AP_CUDA(h, cudaMemcpyAsync(&dr, W.res, sizeof(DResult), cudaMemcpyDeviceToHost, h->stream));
  AP_CUDA(h, cudaMemcpyAsync(x_out, W.xout, (size_t)n * 8, cudaMemcpyDeviceToHost, h->stream));
  if (y_out && P.md > 0) AP_CUDA(h, cudaMemcpyAsync(y_out, W.yout, (size_t)P.md * 8, cudaMemcpyDeviceToHost, h->stream));
  AP_CUDA(h, cudaStreamSynchronize(h->stream));
  if (helper_launched) AP_CUDA(h, cudaStreamSynchronize(h->stream2));     // they leave on the exit tag the main kernel writes last
  if (fused) fused_print_probe(fpl, "fused");
  if (fused && fa.hV > 0 && std::getenv("ADAPROX_HELPER_STATS")) {
    unsigned long long hs[12];
    cudaMemcpy(hs, fa.hsync, sizeof(hs), cudaMemcpyDeviceToHost);
    std::fprintf(stderr, "[adaprox helpers: V=%d R=%d launched=%d roll=%llu abort=%llu chunks by helpers=%llu of %lld (%d chunks per sweep)]\n", fa.hV, fa.hR,
                 (int)helper_launched, hs[0], hs[1], hs[10], (long long)(dr.f_evals) * fa.nchunks, fa.nchunks);
  }
  if (fused && (rc = fused_check(h, fpl))) return rc;
  if (records && dr.n_records > 0)
    AP_CUDA(h, cudaMemcpy(records, W.rec, (size_t)dr.n_records * sizeof(adaprox_record), cudaMemcpyDeviceToHost));
  if (d_ts) {     // phase breakdown of the persistent kernel, averaged over the stamped iterations
    std::vector<unsigned long long> ts((size_t)ts_iters * 8);
    cudaMemcpy(ts.data(), d_ts, ts.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    const int nit = (int)std::min<int64_t>(ts_iters, dr.iters);
    double sum[8] = {0}; int cnt = 0;
    for (int i = 0; i < nit; ++i) {
      const unsigned long long* t = &ts[(size_t)i * 8];
      if (!t[0] || !t[7]) continue;
      for (int k = 0; k < 7; ++k) if (t[k + 1] && t[k]) sum[k] += (double)(t[k + 1] - t[k]) * 1e-3;
      sum[7] += (double)(t[7] - t[0]) * 1e-3;
      ++cnt;
    }
    if (fused) {
      std::fprintf(stderr, "[adaprox fused, us per iteration (pass | rest), CTA 0, clusters=%d C=%d]", fQ, fa.C);
      std::fprintf(stderr, " prologue pass %.0f;", (double)(ts[2] - ts[1]) * 1e-3);
      for (int64_t it = 0; it < ts_iters && it < 12; ++it) {
        const unsigned long long* t = &ts[(size_t)it * 8];
        if (!t[0] || !t[7]) break;
        std::fprintf(stderr, " %.0f|%.0f", (double)(t[3] - t[0]) * 1e-3, (double)(t[7] - t[3]) * 1e-3);
      }
      std::fprintf(stderr, "\n");
      double f3[3] = {0, 0, 0}; int fc = 0;
      for (int i = 1; i < nit; ++i) {                       // stamps 0, 3, 4, 7 of the fused loop; the first iteration is left out
        const unsigned long long* t = &ts[(size_t)i * 8];
        if (!t[0] || !t[3] || !t[4] || !t[7]) continue;
        f3[0] += (double)(t[3] - t[0]) * 1e-3; f3[1] += (double)(t[4] - t[3]) * 1e-3; f3[2] += (double)(t[7] - t[4]) * 1e-3; ++fc;
      }
      if (fc) std::fprintf(stderr, "[adaprox fused, mean of %d iterations, us: sweep + grid barrier %.1f | gradient slice + 4 sums + grid barrier %.1f | stepsize, record, prox, grid barrier %.1f]\n",
                           fc, f3[0] / fc, f3[1] / fc, f3[2] / fc);
    }
    if (cnt && !fused) {
      const char* names_pd[7] = {"P1 F*x (+A*x)", "P2 rows/residual", "P3 F'*r", "P4 grad+reductions", "P5 stepsize/dual", "P6 record/A'y", "P7 prox step"};
      const char* names_gr[7] = {"passes 1+2", "grid barrier 1", "gradient entries", "grid barrier 2", "load gradient + 5 sums", "stepsize + record", "prox step"};
      const char** names = gridres ? names_gr : names_pd;
      std::fprintf(stderr, "[adaprox phase timing] %d iterations, us per iteration:", cnt);
      for (int k = 0; k < 7; ++k) std::fprintf(stderr, " %s=%.1f", names[k], sum[k] / cnt);
      std::fprintf(stderr, " total=%.1f\n", sum[7] / cnt);
    }
  }
  float ms = 0.f;
  cudaEventElapsedTime(&ms, h->ev0, h->ev1);
  std::memset(res, 0, sizeof(*res));
  res->iters = dr.iters; res->flags = dr.flags;
  res->f_evals = dr.f_evals; res->grad_f_evals = dr.grad_f_evals; res->prox_g_evals = dr.prox_g_evals;
  res->prox_h_evals = dr.prox_h_evals; res->A_evals = dr.A_evals; res->At_evals = dr.At_evals;
  res->n_records = dr.n_records;
  res->final_gamma = dr.final_gamma; res->final_sigma = dr.final_sigma; res->final_norm_res = dr.final_norm_res;
  res->solve_ms = ms; res->kernel_launches = h->launches - launches0;
  res->matrix_passes = (P.F.kind == MAT_NONE) ? 0 : (fused ? 1 : (resident ? 3 : (gridres ? 4 : 2)));
  res->collective = (sharded_pd || sharded_fused) ? 2 : 0;
  if ((sharded_pd || sharded_fused) && (rc = p2p_check(h))) return rc;
  return ADAPROX_OK;
}

// ---------------------------------------------------------------------------
// measurement hook
// ---------------------------------------------------------------------------
extern "C" int adaprox_time_kernel(adaprox_handle h, adaprox_id mat, int which, int reps, double* ms_per_pass) {
  if (!h || !ms_per_pass || reps <= 0 || (which != 0 && which != 1)) return fail(h, ADAPROX_ERR_INVALID, "time_kernel: bad arguments");
  HostMatrix* hm;
  int rc = get_mat(h, mat, &hm);
  if (rc) return rc;
  AP_CUDA(h, cudaSetDevice(h->device));
  DMat M = hm->d;
  const int64_t len = which == 0 ? M.n : M.m;
  if ((rc = ws_reset(h, ws_size_doubles(len)))) return rc;
  double* in = ws_doubles(h, len);
  std::vector<double> ones((size_t)len, 1.0);
  AP_CUDA(h, cudaMemcpyAsync(in, ones.data(), (size_t)len * 8, cudaMemcpyHostToDevice, h->stream));
  double* out = nullptr;
  k_gemv_pass<<<h->grid, kThreads, kRingBytes, h->stream>>>(M, which, in, out);   // warm-up
  h->launches++;
  AP_CUDA(h, cudaEventRecord(h->ev0, h->stream));
  for (int r = 0; r < reps; ++r) { k_gemv_pass<<<h->grid, kThreads, kRingBytes, h->stream>>>(M, which, in, out); h->launches++; }
  AP_CUDA(h, cudaEventRecord(h->ev1, h->stream));
  AP_CUDA(h, cudaStreamSynchronize(h->stream));
  AP_CUDA(h, cudaGetLastError());
  float ms = 0.f;
  AP_CUDA(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  *ms_per_pass = (double)ms / reps;
  return ADAPROX_OK;
}

#include "comm.inl"
#include "generate.inl"
#include "path.inl"
#include "hessian.inl"
