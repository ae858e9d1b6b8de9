// comm.inl -- included at the end of api.cu (same translation unit).
//
// Row-sharded AdaPGM: rows of A (and b / labels) are split
// in contiguous blocks over the ranks, x / grad / stepsize state are replicated.
// Per iteration every rank sweeps its shard (once with the fused kernel on dense
// least squares, twice -- A*x, then A'r -- otherwise) and ONE all-reduce of n+2
// doubles combines the A'r partials with the value sums.  Every rank obtains
// bit-identical sums, so the replicated prox / stepsize arithmetic stays in lock
// step with no second collective.  The all-reduce is either ncclAllReduce between
// ordinary launches on one stream (three launches per iteration with the fused
// sweep, six on the two-pass path) or, when the peer exchange blocks are attached
// (adaprox_p2p_*), done INSIDE the sweep kernel over NVLink peer memory (p2p.cuh).
// Convergence is a device flag polled by the host once per batch of iterations, so
// there is no host round trip per iteration.  This file also owns the NCCL loader,
// the CUDA-IPC exchange blocks and their C entry points; the row-sharded
// primal-dual solves run in the persistent kernel of solver_pd.cuh.
#include <dlfcn.h>
#include <nccl.h>

namespace adaprox {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  if (api.lib) return &api;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) { api.err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return &api; }
  api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
  api.AllReduce = (decltype(api.AllReduce))dlsym(api.lib, "ncclAllReduce");
  api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
  if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce || !api.GetErrorString) {
    api.err = "libnccl is missing a required symbol";
    dlclose(api.lib);
    api.lib = nullptr;
  }
  return &api;
}

// Peer-mapped exchange block of the in-kernel all-reduce (solver_fused.cuh, sweep-only mode):
//   [ flags: kP2PMaxRanks x u64, padded to 1 KB ][ buffer 0: cap doubles ][ buffer 1: cap doubles ]
// allocated with cudaMalloc by every rank, exported with cudaIpcGetMemHandle and opened by the peers.
struct P2P {
  bool attached = false;
  int nranks = 1, rank = 0;
  int64_t cap = 0;                               // doubles per buffer
  char* local = nullptr;                         // this rank's block
  char* peer[kP2PMaxRanks] = {};                 // every rank's block as mapped here (peer[rank] == local)
  int* err = nullptr;                            // device flag: a peer did not show up
};

struct Comm {
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
  P2P p2p;
};

static constexpr size_t kP2PFlagBytes = 1024;
static double* p2p_buf(char* block, int64_t cap, int which) { return reinterpret_cast<double*>(block + kP2PFlagBytes + (size_t)which * (size_t)cap * 8); }

// kernel-side view of the exchange blocks (p2p.cuh)
void p2p_fill(adaprox_ctx* h, P2PArgs* pa) {
  *pa = P2PArgs{};
  if (!h->comm || !h->comm->p2p.attached) return;
  const P2P& pp = h->comm->p2p;
  pa->n = pp.nranks; pa->rank = pp.rank; pa->cap = pp.cap; pa->err = pp.err;
  for (int q = 0; q < pp.nranks; ++q) {
    pa->buf[q][0] = p2p_buf(pp.peer[q], pp.cap, 0);
    pa->buf[q][1] = p2p_buf(pp.peer[q], pp.cap, 1);
    pa->flags[q] = reinterpret_cast<unsigned long long*>(pp.peer[q]);
  }
}
int comm_nranks(adaprox_ctx* h) { return h->comm ? h->comm->nranks : 1; }
bool p2p_ready(adaprox_ctx* h, int64_t count) { return h->comm && h->comm->p2p.attached && h->comm->p2p.cap >= count; }
int p2p_check(adaprox_ctx* h) {
  if (!h->comm || !h->comm->p2p.err) return ADAPROX_OK;
  int perr = 0;
  AP_CUDA(h, cudaMemcpy(&perr, h->comm->p2p.err, sizeof(int), cudaMemcpyDeviceToHost));
  if (perr) return fail(h, ADAPROX_ERR_COMM, "in-kernel all-reduce: a peer rank did not arrive within 5 s (call adaprox_p2p_reset on every rank before the next sharded solve)");
  return ADAPROX_OK;
}

int comm_allreduce_sum(adaprox_ctx* h, double* buf_dev, int64_t count) {
  if (!h->comm) return fail(h, ADAPROX_ERR_COMM, "no communicator attached");
  NcclApi* api = nccl_api();
  ncclResult_t r = api->AllReduce(buf_dev, buf_dev, (size_t)count, ncclDouble, ncclSum, h->comm->comm, h->stream);
  if (r != ncclSuccess) return fail(h, ADAPROX_ERR_COMM, std::string("ncclAllReduce: ") + api->GetErrorString(r));
  return ADAPROX_OK;
}

void comm_destroy(adaprox_ctx* h) {
  if (h->comm) {
    P2P& pp = h->comm->p2p;
    for (int q = 0; q < kP2PMaxRanks; ++q)
      if (pp.peer[q] && pp.peer[q] != pp.local) cudaIpcCloseMemHandle(pp.peer[q]);
    if (pp.local) cudaFree(pp.local);
    if (pp.err) cudaFree(pp.err);
    pp = P2P{};
    NcclApi* api = nccl_api();
    if (api->lib && h->comm->comm) api->CommDestroy(h->comm->comm);
    delete h->comm;
    h->comm = nullptr;
  }
}

// ---------------------------------------------------------------------------
// split-phase kernels of the sharded AdaPGM iteration
// ---------------------------------------------------------------------------
struct ShState {
  double gamma, sigma, s0, s1, norm_res;
  long long it;            // on `done`: the iteration at which the loop stopped
  int done;
  unsigned flags;
  long long n_eval, n_grad, n_proxg, n_rec;
};

struct ShArgs {
  DProblem P; DOpts O; DWork W;
  ShState* st;             // [2]: the kernels of iteration `it` read st[it & 1], k_sh_F writes st[(it + 1) & 1]
  long long it;            // iteration being enqueued (host-side counter; 0 = prologue)
  double* gbuf;            // [n + 2]: A'r sums, value sums  (all-reduced in place)
};

__device__ __forceinline__ const ShState& sh_state(const ShArgs& a, long long it) { return a.st[it & 1]; }
__device__ __forceinline__ long long sh_current(const ShArgs& a, bool& done) {
  done = a.st[a.it & 1].done != 0;
  return a.it;
}

__global__ void __launch_bounds__(kThreads, 2) k_sh_A(ShArgs a) {
  extern __shared__ __align__(1024) unsigned char dyn_smem[];
  __shared__ double s_scr[kWarps * 8 + kMaxRed];
  __shared__ double s_part[kPartRows * kWarps];
  __shared__ unsigned long long s_bars[2 * kStages];
  Sh sh;
  sh_init(sh, dyn_smem, s_scr, s_part, s_bars);
  bool done; const long long it = sh_current(a, done);
  if (done) return;
  f_phase_A(a.P, a.W, a.W.xb[it % 3], sh, s_scr, blockIdx.x, gridDim.x);
}
__global__ void __launch_bounds__(kThreads, 2) k_sh_B(ShArgs a) {
  __shared__ double s_scr[kWarps * 8 + kMaxRed];
  bool done; const long long it = sh_current(a, done);
  if (done) return;
  f_phase_B(a.P, a.W, a.W.xb[it % 3], s_scr, blockIdx.x, gridDim.x);
}
__global__ void __launch_bounds__(kThreads, 2) k_sh_C(ShArgs a) {
  extern __shared__ __align__(1024) unsigned char dyn_smem[];
  __shared__ double s_scr[kWarps * 8 + kMaxRed];
  __shared__ double s_part[kPartRows * kWarps];
  __shared__ unsigned long long s_bars[2 * kStages];
  Sh sh;
  sh_init(sh, dyn_smem, s_scr, s_part, s_bars);
  bool done; sh_current(a, done);
  if (done) return;
  f_phase_C(a.P, a.W, sh, blockIdx.x, gridDim.x);
}
// local partial sums -> the all-reduce buffer
__global__ void __launch_bounds__(kThreads, 2) k_sh_D(ShArgs a) {
  __shared__ double s_scr[kWarps * 8 + kMaxRed];
  bool done; sh_current(a, done);
  if (done) return;
  const int b = blockIdx.x, G = gridDim.x;
  const int64_t nmat = a.P.F.n;
  int64_t j0, j1;
  cta_slice(nmat, b, G, j0, j1);
  gsum_slice(a.P.F, j0, j1, a.gbuf, G);
  double tot[2];
  grid_totals<2>(a.W.red, G, SLOT_F0, tot, s_scr);
  if (b == 0 && threadIdx.x == 0) { a.gbuf[a.P.n] = tot[0]; a.gbuf[a.P.n + 1] = tot[1]; }
}
// Block/grid reduction policy of the split-phase kernels (256-thread CTAs).
struct Red256 {
  static constexpr int NT = kThreads;
  double* scr;
  template <int K> __device__ __forceinline__ void store(double (&v)[K], double* red, int G, int slot0) { block_reduce_store<K>(v, red, G, slot0, scr); }
  template <int K> __device__ __forceinline__ void totals(const double* red, int G, int slot0, double (&out)[K]) { grid_totals<K>(red, G, slot0, out, scr); }
};

// after the all-reduce: gradient, primal residual, stepsize reductions (P4)
template <class R>
__device__ __forceinline__ void sh_phase_E(const ShArgs& a, long long it, R red) {
  const DProblem& P = a.P;
  const int b = blockIdx.x, G = gridDim.x;
  const ShState& st = sh_state(a, it);
  int64_t j0, j1;
  cta_slice(P.n, b, G, j0, j1);
  double* grad = a.W.gb[it & 1];
  const double* grad_prev = a.W.gb[(it + 1) & 1];
  const double* x = a.W.xb[it % 3];
  const double* x_prev = a.W.xb[(it + 2) % 3];
  const double f1 = a.gbuf[P.n + 1];
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int64_t j = j0 + threadIdx.x; j < j1; j += R::NT) {
    double gj = a.gbuf[j];
    if (P.f_kind == ADAPROX_F_LOGISTIC) gj = (j == P.n - 1) ? f1 / P.f_N : gj / P.f_N;
    grad[j] = gj;
    if (it > 0) {
      const double xj = x[j];
      const double pr = (a.W.v[j] - xj) / st.gamma + gj;
      const double dg = gj - grad_prev[j], dx = xj - x_prev[j];
      acc[0] = fma(pr, pr, acc[0]);
      acc[1] = fma(dg, dg, acc[1]);
      acc[2] = fma(dg, dx, acc[2]);
      acc[3] = fma(dx, dx, acc[3]);
    }
  }
  red.template store<4>(acc, a.W.red, G, SLOT_PR);
}
// stepsize, residual, record, convergence, prox step (P5 + P7); advances the state
template <class R>
__device__ __forceinline__ void sh_phase_F(const ShArgs& a, long long it, R red) {
  const DProblem& P = a.P;
  const DOpts& O = a.O;
  const int b = blockIdx.x, G = gridDim.x;
  const ShState st = sh_state(a, it);
  ShState nx = st;
  int64_t j0, j1;
  cta_slice(P.n, b, G, j0, j1);
  double gamma = st.gamma, sigma = st.sigma, s0 = st.s0, s1 = st.s1;
  bool stop = false;
  if (it > 0) {
    double t4[4], tg[1] = {0.0};
    red.template totals<4>(a.W.red, G, SLOT_PR, t4);
    if (O.want_objective) red.template totals<1>(a.W.red, G, gval_slot(it), tg);
    const double gamma_prev = gamma;
    rule_step(O, t4[1], t4[2], t4[3], gamma, sigma, s0, s1);
    const double norm_res = sqrt(norm_sq_jl(t4[0]) + adapgm_dual_res_sq(gamma, gamma_prev, sigma));   // :348 (dual part: 0, or NaN -- phases.cuh)
    nx.norm_res = norm_res;
    if (!(gamma == gamma) || !(norm_res == norm_res) || isinf(gamma)) nx.flags |= ADAPROX_FLAG_NONFINITE;
    if (b == 0 && threadIdx.x == 0 && a.W.rec != nullptr && it <= O.max_records) {
      adaprox_record rc;
      rc.it = it; rc.gamma = gamma; rc.sigma = sigma; rc.norm_res = norm_res;
      rc.f_x = f_value(P, a.gbuf[P.n], a.gbuf[P.n + 1], 0.0);
      rc.g_x = O.want_objective ? prox_value_finish(P.g.kind, P.g.lambda, tg[0]) : NAN;
      rc.h_Ax = O.want_objective ? 0.0 : NAN;
      rc.f_evals = st.n_eval; rc.grad_f_evals = st.n_grad; rc.prox_g_evals = st.n_proxg; rc.prox_h_evals = 0;
      rc.A_evals = 0; rc.At_evals = 0;
      a.W.rec[it - 1] = rc;
    }
    if (it <= O.max_records) nx.n_rec = it;
    if (norm_res <= O.tol) { stop = true; nx.flags |= ADAPROX_FLAG_CONVERGED; }
  }
  nx.gamma = gamma; nx.sigma = sigma; nx.s0 = s0; nx.s1 = s1;
  if (!stop) {
    const double* x = a.W.xb[it % 3];
    const double* grad = a.W.gb[it & 1];
    double* xn = a.W.xb[(it + 1) % 3];
    double acc[1] = {0.0};
    for (int64_t j = j0 + threadIdx.x; j < j1; j += R::NT) {
      const double vj = x[j] - gamma * grad[j];
      a.W.v[j] = vj;
      const double xj = prox_elem(P.g, vj, gamma, j, 0.0);
      xn[j] = xj;
      if (O.want_objective) acc[0] += prox_value_elem(P.g, xj, j);
    }
    red.template store<1>(acc, a.W.red, G, gval_slot(it + 1));
    nx.n_proxg = st.n_proxg + 1;
    if (it >= O.maxit) { nx.done = 1; nx.it = it; }           // maxit reached: x is the last prox, xb[(it + 1) % 3]
    else { nx.n_eval = st.n_eval + 1; nx.n_grad = st.n_grad + 1; }
  } else {
    nx.done = 1; nx.it = it;                                  // converged: x is xb[it % 3]
  }
  if (b == 0 && threadIdx.x == 0) a.st[(it + 1) & 1] = nx;
}
__global__ void __launch_bounds__(kThreads, 2) k_sh_E(ShArgs a) {
  __shared__ double s_scr[kWarps * 8 + kMaxRed];
  bool done; const long long it = sh_current(a, done);
  if (done) return;
  sh_phase_E(a, it, Red256{s_scr});
}
__global__ void __launch_bounds__(kThreads, 2) k_sh_F(ShArgs a) {
  __shared__ double s_scr[kWarps * 8 + kMaxRed];
  bool done; const long long it = sh_current(a, done);
  if (done) {          // keep the final state alive in both slots
    if (blockIdx.x == 0 && threadIdx.x == 0) a.st[(it + 1) & 1] = a.st[it & 1];
    return;
  }
  sh_phase_F(a, it, Red256{s_scr});
}

// Row-sharded AdaPGM on a dense least-squares term: the four launches k_sh_A .. k_sh_D are replaced by ONE launch
// of k_adapgm_fused in sweep-only mode (solver_fused.cuh), which sweeps this rank's row shard once for x_it and
// leaves the shard's A'r partial and value sum in gbuf for the all-reduce; k_sh_E / k_sh_F finish the iteration.

int comm_setup_kernels() {
  if (cudaFuncSetAttribute((const void*)k_sh_A, cudaFuncAttributeMaxDynamicSharedMemorySize, kRingBytes) != cudaSuccess) return -1;
  if (cudaFuncSetAttribute((const void*)k_sh_C, cudaFuncAttributeMaxDynamicSharedMemorySize, kRingBytes) != cudaSuccess) return -1;
  return 0;
}

int solve_sharded(adaprox_ctx* h, const adaprox_problem* p, const adaprox_options* o, const DProblem& P, const DOpts& Oin,
                  HostMatrix* fm, HostMatrix* am, const double* x0, const double* y0, double* x_out, double* y_out,
                  adaprox_record* records, adaprox_result* res) {
  (void)p; (void)y0; (void)y_out; (void)am;
  if (!h->comm) return fail(h, ADAPROX_ERR_COMM, "matrix is a row shard but no communicator is attached (adaprox_comm_init)");
  if (o->solver != ADAPROX_S_ADAPTIVE_PROXGRAD)
    return fail(h, ADAPROX_ERR_UNSUPPORTED, "row-sharded solves support adaptive_proxgrad / fixed_proxgrad only");
  if (P.f_kind != ADAPROX_F_LEAST_SQUARES && P.f_kind != ADAPROX_F_LOGISTIC)
    return fail(h, ADAPROX_ERR_UNSUPPORTED, "row-sharded solves support least-squares and logistic smooth terms only");
  if (!fm || fm->d.kind == MAT_NONE) return fail(h, ADAPROX_ERR_INVALID, "sharded solve without a matrix");
  if (P.g.kind == ADAPROX_P_NORM_L2 || P.g.conjugate)
    return fail(h, ADAPROX_ERR_UNSUPPORTED, "row-sharded AdaPGM: g = NormL2 / conjugate g is not supported (separable g only)");
  DOpts O = Oin;
  const int64_t n = P.n, mf = P.F.m;
  const int64_t nrec = std::min<int64_t>(O.max_records, O.maxit);
  O.max_records = nrec;
  int G = h->grid;
  // dense least squares: the single-pass fused kernel in sweep-only mode replaces k_sh_A .. k_sh_D
  bool fused = fused_eligible(o, P, (fm->m_global + h->comm->nranks - 1) / h->comm->nranks);
  FusedPlan fpl;
  if (fused) {
    int frc = fused_plan(h, (const void*)k_adapgm_fused, P, false, &fpl);
    if (frc < 0) return frc;
    if (frc == 1) fused = false; else G = fpl.G;
  }
  FusedArgs& fa = fpl.fa;
  const int Gred = std::max(G, h->grid);             // k_sh_E / k_sh_F always run on the regular grid
  size_t need = (fused ? fused_ws_bytes(fpl) : 0) + 7 * ws_size_doubles(n) + ws_size_doubles(n + 2) + ws_size_doubles(mf) + ws_size_doubles((int64_t)kMaxRed * Gred) +
                ws_size_doubles((nrec * (int64_t)sizeof(adaprox_record) + 7) / 8) + ws_size_doubles(2 * (sizeof(ShState) + 7) / 8);
  int rc;
  if ((rc = ws_reset(h, need))) return rc;
  ShArgs a{};
  a.P = P; a.O = O;
  DWork& W = a.W;
  for (int k = 0; k < 3; ++k) W.xb[k] = ws_doubles(h, n);
  for (int k = 0; k < 2; ++k) W.gb[k] = ws_doubles(h, n);
  W.v = ws_doubles(h, n);
  W.xout = ws_doubles(h, n);
  a.gbuf = ws_doubles(h, n + 2);
  W.r = ws_doubles(h, mf);
  W.red = ws_doubles(h, (int64_t)kMaxRed * Gred);
  W.rec = nrec > 0 ? reinterpret_cast<adaprox_record*>(ws_doubles(h, (nrec * (int64_t)sizeof(adaprox_record) + 7) / 8)) : nullptr;
  a.st = reinterpret_cast<ShState*>(ws_doubles(h, 2 * (sizeof(ShState) + 7) / 8));
  if (fused && (rc = fused_ws_alloc(h, &fpl))) return rc;

  ShState init[2];
  std::memset(init, 0, sizeof(init));
  rule_init(O, init[0].gamma, init[0].sigma, init[0].s0, init[0].s1);
  init[0].it = 0; init[0].n_eval = 1; init[0].n_grad = 1; init[0].norm_res = INFINITY;
  init[1] = init[0];
  AP_CUDA(h, cudaMemcpyAsync(a.st, init, sizeof(init), cudaMemcpyHostToDevice, h->stream));
  AP_CUDA(h, cudaMemcpyAsync(W.xb[0], x0, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
  AP_CUDA(h, cudaMemsetAsync(W.red, 0, (size_t)kMaxRed * Gred * 8, h->stream));
  AP_CUDA(h, cudaMemsetAsync(a.gbuf, 0, (size_t)(n + 2) * 8, h->stream));
  const int64_t launches0 = h->launches;
  AP_CUDA(h, cudaEventRecord(h->ev0, h->stream));

  auto one_iteration = [&](int64_t it) -> int {
    a.it = it;
    k_sh_A<<<G, kThreads, kRingBytes, h->stream>>>(a);
    k_sh_B<<<G, kThreads, 0, h->stream>>>(a);
    k_sh_C<<<G, kThreads, kRingBytes, h->stream>>>(a);
    k_sh_D<<<G, kThreads, 0, h->stream>>>(a);
    int r = comm_allreduce_sum(h, a.gbuf, n + 2);
    if (r) return r;
    k_sh_E<<<G, kThreads, 0, h->stream>>>(a);
    k_sh_F<<<G, kThreads, 0, h->stream>>>(a);
    h->launches += 6;
    return ADAPROX_OK;
  };

  // fused path: launch L = finish iteration L-1 + sweep for x_L; launches 0 .. maxit+1 (the last one only finishes)
  // in-kernel all-reduce over NVLink peer memory when the exchange blocks are attached and large enough (p2p.cuh)
  const bool use_p2p = fused && h->comm->p2p.attached && h->comm->p2p.cap >= n + 2 && !std::getenv("ADAPROX_NO_P2P");
  // fused path: one sweep kernel instead of k_sh_A .. k_sh_D
  auto one_iteration_fused = [&](int64_t it) -> int {
    a.it = it;
    fa.bar_base = (unsigned long long)(use_p2p ? 4 : 1) * (unsigned long long)G * (unsigned long long)it;   // grid barriers per launch
    fa.next_base = (unsigned long long)(fa.nchunks + fpl.Q) * (unsigned long long)it;
    fa.sweep_only = 1;
    fa.sh_x = a.W.xb[it % 3];
    fa.sh_gbuf = a.gbuf;
    fa.sh_done = &a.st[it & 1].done;
    if (use_p2p) p2p_fill(h, &fa.p2p);
    void* fargs[] = {&a.P, &a.O, &a.W, &fa};
    cudaError_t e = cudaLaunchKernelExC(&fpl.cfg, (const void*)k_adapgm_fused, fargs);
    if (e != cudaSuccess) return fail(h, ADAPROX_ERR_CUDA, std::string("fused cluster launch (sharded): ") + cudaGetErrorString(e));
    if (!use_p2p) { int r = comm_allreduce_sum(h, a.gbuf, n + 2); if (r) return r; }    // else: reduced inside the kernel
    k_sh_E<<<h->grid, kThreads, 0, h->stream>>>(a);
    k_sh_F<<<h->grid, kThreads, 0, h->stream>>>(a);
    h->launches += 3;
    return ADAPROX_OK;
  };

  ShState live[2];
  int64_t enqueued = 0;               // gradient evaluations enqueued (prologue + iterations)
  const int64_t total = O.maxit + 1;
  bool finished = false;
  while (!finished) {
    const int64_t batch = std::min<int64_t>(32, total - enqueued);
    for (int64_t k = 0; k < batch; ++k) if ((rc = fused ? one_iteration_fused(enqueued + k) : one_iteration(enqueued + k))) return rc;
    enqueued += batch;
    AP_CUDA(h, cudaMemcpyAsync(live, a.st, sizeof(live), cudaMemcpyDeviceToHost, h->stream));
    AP_CUDA(h, cudaStreamSynchronize(h->stream));
    AP_CUDA(h, cudaGetLastError());
    if (use_p2p && (rc = p2p_check(h))) return rc;            // a peer was lost: the kernels no longer wait, stop enqueueing
    if (live[enqueued & 1].done || enqueued >= total) finished = true;
  }
  const int cur_idx = (int)(enqueued & 1);          // written by the last k_sh_F
  AP_CUDA(h, cudaEventRecord(h->ev1, h->stream));
  if (use_p2p && (rc = p2p_check(h))) return rc;
  if (fused) fused_print_probe(fpl, "sharded fused");
  if (fused && (rc = fused_check(h, fpl))) return rc;
  const ShState& cur = live[cur_idx];
  const bool converged = (cur.flags & ADAPROX_FLAG_CONVERGED) != 0;
  // converged at iteration k: x is xb[k % 3]; maxit: the last prox, xb[(maxit + 1) % 3]
  const int64_t it_final = converged ? cur.it : O.maxit;
  const double* xres = converged ? W.xb[it_final % 3] : W.xb[(O.maxit + 1) % 3];
  AP_CUDA(h, cudaMemcpyAsync(x_out, xres, (size_t)n * 8, cudaMemcpyDeviceToHost, h->stream));
  AP_CUDA(h, cudaStreamSynchronize(h->stream));
  if (records && cur.n_rec > 0)
    AP_CUDA(h, cudaMemcpy(records, W.rec, (size_t)cur.n_rec * sizeof(adaprox_record), cudaMemcpyDeviceToHost));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, h->ev0, h->ev1);
  std::memset(res, 0, sizeof(*res));
  res->iters = it_final; res->flags = cur.flags;
  res->f_evals = cur.n_eval; res->grad_f_evals = cur.n_grad; res->prox_g_evals = cur.n_proxg;
  res->n_records = cur.n_rec;
  res->final_gamma = cur.gamma; res->final_sigma = cur.sigma; res->final_norm_res = cur.norm_res;
  res->solve_ms = ms; res->kernel_launches = h->launches - launches0;
  res->matrix_passes = fused ? 1 : 2;
  res->collective = use_p2p ? 2 : 1;
  return ADAPROX_OK;
}

}  // namespace adaprox

extern "C" int adaprox_comm_unique_id(void* id128) {
  if (!id128) return ADAPROX_ERR_INVALID;
  adaprox::NcclApi* api = adaprox::nccl_api();
  if (!api->lib) return ADAPROX_ERR_COMM;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  if (api->GetUniqueId(&id) != ncclSuccess) return ADAPROX_ERR_COMM;
  std::memcpy(id128, &id, 128);
  return ADAPROX_OK;
}

extern "C" int adaprox_comm_init(adaprox_handle h, int nranks, int rank, const void* id128) {
  if (!h || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return adaprox::fail(h, ADAPROX_ERR_INVALID, "comm_init: bad arguments");
  adaprox::NcclApi* api = adaprox::nccl_api();
  if (!api->lib) return adaprox::fail(h, ADAPROX_ERR_COMM, api->err);
  AP_CUDA(h, cudaSetDevice(h->device));
  adaprox::comm_destroy(h);
  ncclUniqueId id;
  std::memcpy(&id, id128, 128);
  adaprox::Comm* c = new adaprox::Comm();
  c->nranks = nranks; c->rank = rank;
  ncclResult_t r = api->CommInitRank(&c->comm, nranks, id, rank);
  if (r != ncclSuccess) { delete c; return adaprox::fail(h, ADAPROX_ERR_COMM, std::string("ncclCommInitRank: ") + api->GetErrorString(r)); }
  h->comm = c;
  return ADAPROX_OK;
}

extern "C" int adaprox_p2p_export(adaprox_handle h, int64_t n_max, void* ipc_handle64) {
  if (!h || !ipc_handle64 || n_max < 1) return adaprox::fail(h, ADAPROX_ERR_INVALID, "p2p_export: bad arguments");
  if (!h->comm) return adaprox::fail(h, ADAPROX_ERR_COMM, "p2p_export: call adaprox_comm_init first");
  AP_CUDA(h, cudaSetDevice(h->device));
  adaprox::P2P& pp = h->comm->p2p;
  if (pp.local) return adaprox::fail(h, ADAPROX_ERR_INVALID, "p2p_export: already exported on this handle");
  pp.cap = (n_max + 2 + 15) / 16 * 16;
  const size_t bytes = adaprox::kP2PFlagBytes + 2 * (size_t)pp.cap * 8;
  AP_CUDA(h, cudaMalloc(&pp.local, bytes));
  AP_CUDA(h, cudaMemset(pp.local, 0, bytes));
  AP_CUDA(h, cudaMalloc(&pp.err, sizeof(int)));
  AP_CUDA(h, cudaMemset(pp.err, 0, sizeof(int)));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t hd;
  AP_CUDA(h, cudaIpcGetMemHandle(&hd, pp.local));
  std::memcpy(ipc_handle64, &hd, 64);
  return ADAPROX_OK;
}

extern "C" int adaprox_p2p_attach(adaprox_handle h, int nranks, int rank, const void* ipc_handles) {
  if (!h || !ipc_handles || nranks < 1 || nranks > adaprox::kP2PMaxRanks || rank < 0 || rank >= nranks)
    return adaprox::fail(h, ADAPROX_ERR_INVALID, "p2p_attach: bad arguments (at most 8 ranks)");
  if (!h->comm || !h->comm->p2p.local) return adaprox::fail(h, ADAPROX_ERR_COMM, "p2p_attach: call adaprox_p2p_export first");
  if (h->comm->nranks != nranks || h->comm->rank != rank) return adaprox::fail(h, ADAPROX_ERR_INVALID, "p2p_attach: ranks differ from adaprox_comm_init");
  AP_CUDA(h, cudaSetDevice(h->device));
  adaprox::P2P& pp = h->comm->p2p;
  for (int q = 0; q < nranks; ++q) {
    if (q == rank) { pp.peer[q] = pp.local; continue; }
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, static_cast<const char*>(ipc_handles) + 64 * q, 64);
    void* p = nullptr;
    AP_CUDA(h, cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
    pp.peer[q] = static_cast<char*>(p);
  }
  pp.nranks = nranks; pp.rank = rank; pp.attached = true;
  return ADAPROX_OK;
}

extern "C" int adaprox_p2p_reset(adaprox_handle h) {
  if (!h) return ADAPROX_ERR_INVALID;
  if (!h->comm || !h->comm->p2p.local) return adaprox::fail(h, ADAPROX_ERR_COMM, "p2p_reset: no exchange block (adaprox_p2p_export)");
  AP_CUDA(h, cudaSetDevice(h->device));
  AP_CUDA(h, cudaStreamSynchronize(h->stream));
  AP_CUDA(h, cudaMemset(h->comm->p2p.local, 0, adaprox::kP2PFlagBytes));
  AP_CUDA(h, cudaMemset(h->comm->p2p.err, 0, sizeof(int)));
  return ADAPROX_OK;
}

extern "C" int adaprox_comm_info(adaprox_handle h, int* nranks, int* rank) {
  if (!h) return ADAPROX_ERR_INVALID;
  if (nranks) *nranks = h->comm ? h->comm->nranks : 1;
  if (rank) *rank = h->comm ? h->comm->rank : 0;
  return ADAPROX_OK;
}
