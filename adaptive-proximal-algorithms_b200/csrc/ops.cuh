// ops.cuh -- one-call operator kernels (A*x, A'*y, f + gradient, prox) and the
// synthetic-data kernels.  The operator kernels reuse the solver's grid phases
// verbatim, so the parity tests that go through adaprox_mul / adaprox_amul /
// adaprox_eval_f / adaprox_prox_eval exercise exactly the code the persistent
// solver kernels run.
#pragma once
#include "phases_pre.cuh"

namespace adaprox {

enum { OP_MUL = 0, OP_AMUL = 1, OP_EVALF = 2, OP_PROX = 3 };

struct OpArgs {
  int op;
  DMat M;                 // MUL / AMUL
  DProblem P;             // EVALF
  DProx px; double gamma; int64_t len;   // PROX
  const double* in;       // x / y
  double* out;            // result vector (may be null for EVALF without gradient)
  double* scal;           // [4] scalar results
  int want_grad;
};

__global__ void __launch_bounds__(kThreads, 2) k_ops(OpArgs a, DWork W) {
  cg::grid_group grid = cg::this_grid();
  const int b = blockIdx.x, G = gridDim.x;
  extern __shared__ __align__(1024) unsigned char dyn_smem[];
  __shared__ double s_scr[kWarps * 8 + kMaxRed];
  __shared__ double s_part[kPartRows * kWarps];
  __shared__ unsigned long long s_bars[2 * kStages];
  Sh sh;
  sh_init(sh, dyn_smem, s_scr, s_part, s_bars);
  const int64_t tid = (int64_t)b * kThreads + threadIdx.x, nt = (int64_t)G * kThreads;

  if (a.op == OP_MUL) {
    gemv_n_phase(a.M, a.in, sh, b, G);
    grid.sync();
    for (int64_t i = tid; i < a.M.m; i += nt) a.out[i] = zsum(a.M, i);
  } else if (a.op == OP_AMUL) {
    gemv_t_phase(a.M, a.in, sh, b, G);
    grid.sync();
    int64_t j0, j1;
    cta_slice(a.M.n, b, G, j0, j1);
    gsum_slice(a.M, j0, j1, a.out, G);
  } else if (a.op == OP_EVALF) {
    const DProblem& P = a.P;
    f_phase_pre(grid, P, W, a.in, sh, b, G, nullptr);
    f_phase_A(P, W, a.in, sh, s_scr, b, G);
    grid.sync();
    f_phase_B(P, W, a.in, s_scr, b, G);
    grid.sync();
    if (a.want_grad) f_phase_C(P, W, sh, b, G);
    grid.sync();
    double tot[2], xx[1] = {0.0};
    grid_totals<2>(W.red, G, SLOT_F0, tot, s_scr);
    if (P.f_kind == ADAPROX_F_CUBIC) grid_totals<1>(W.red, G, SLOT_XX0, xx, s_scr);
    if (tid == 0) a.scal[0] = f_value(P, tot[0], tot[1], xx[0]);
    if (a.want_grad) {
      int64_t j0, j1;
      cta_slice(P.n, b, G, j0, j1);
      grad_slice(P, W, j0, j1, a.out, tot[1], G);
    }
  } else {   // OP_PROX: ProximalCore.prox(f, x, gamma) -> (y, f(y))
    double l2scale = 0.0;
    if (a.px.kind == ADAPROX_P_NORM_L2) {
      double acc[1] = {0.0};
      for (int64_t i = tid; i < a.len; i += nt) {
        const double z = prox_l2_arg(a.px, a.px.conjugate ? a.in[i] / a.gamma : a.in[i], i);
        acc[0] = fma(z, z, acc[0]);
      }
      block_reduce_store<1>(acc, W.red, G, SLOT_L2, s_scr);
      grid.sync();
      double t[1];
      grid_totals<1>(W.red, G, SLOT_L2, t, s_scr);
      l2scale = prox_l2_scale(a.px.lambda, a.px.conjugate ? 1.0 / a.gamma : a.gamma, t[0]);
    }
    double acc[1] = {0.0};
    for (int64_t i = tid; i < a.len; i += nt) {
      const double yi = a.px.conjugate ? prox_conj_elem(a.px, a.in[i], a.gamma, i, l2scale)
                                       : prox_elem(a.px, a.in[i], a.gamma, i, l2scale);
      a.out[i] = yi;
      acc[0] += prox_value_elem(a.px, yi, i);
    }
    block_reduce_store<1>(acc, W.red, G, SLOT_HVAL, s_scr);
    grid.sync();
    double t[1];
    grid_totals<1>(W.red, G, SLOT_HVAL, t, s_scr);
    // the value of a conjugate is not used by any solver (`y, _ = prox(...)`); report f(y) of the base function
    if (tid == 0) a.scal[0] = prox_value_finish(a.px.kind, a.px.lambda, t[0]);
  }
}

// One pass of a matrix kernel family, for adaprox_time_kernel (roofline measurement).
__global__ void __launch_bounds__(kThreads, 2) k_gemv_pass(DMat M, int which, const double* in, double* out) {
  const int b = blockIdx.x, G = gridDim.x;
  extern __shared__ __align__(1024) unsigned char dyn_smem[];
  __shared__ double s_scr[kWarps * 8 + kMaxRed];
  __shared__ double s_part[kPartRows * kWarps];
  __shared__ unsigned long long s_bars[2 * kStages];
  Sh sh;
  sh_init(sh, dyn_smem, s_scr, s_part, s_bars);
  if (which == 0) gemv_n_phase(M, in, sh, b, G);
  else gemv_t_phase(M, in, sh, b, G);
}

// ---------------------------------------------------------------------------
// counter-based RNG: SplitMix64 addressed by (key, index); identical to
// synth.py (bits64 / uniform01)
// ---------------------------------------------------------------------------
__host__ __device__ inline uint64_t sm64_mix(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ inline uint64_t stream_key(uint64_t seed, uint64_t stream) {
  return sm64_mix(sm64_mix(seed) ^ ((stream + 1) * 0x9E3779B97F4A7C15ull));
}
__host__ __device__ inline double uniform01(uint64_t key, uint64_t index) {
  const uint64_t z = sm64_mix(key + (index + 1) * 0x9E3779B97F4A7C15ull);
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

// rows [row0, row0+m) of `rand(M, n) .* 2 .- 1` (lasso/runme.jl:50), row-major padded
__global__ void k_fill_uniform_pm1(double* a, int64_t m, int64_t n, int64_t ld, int64_t row0, uint64_t key) {
  const int64_t total = m * ld;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / ld, j = e - i * ld;
    a[e] = (j < n) ? uniform01(key, (uint64_t)((row0 + i) * n + j)) * 2.0 - 1.0 : 0.0;
  }
}

// A = C * diagm(alpha)  (lasso/runme.jl:69)
__global__ void k_scale_columns(double* a, int64_t m, int64_t n, int64_t ld, const double* alpha) {
  const int64_t total = m * ld;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = e % ld;
    if (j < n) a[e] *= alpha[j];
  }
}

// column-major (Julia) -> padded row-major, through a shared-memory tile
__global__ void k_colmajor_to_rowmajor(const double* __restrict__ src, int64_t m, int64_t n, int64_t lda, double* dst, int64_t ld) {
  __shared__ double tile[32][33];
  const int64_t i0 = (int64_t)blockIdx.x * 32, j0 = (int64_t)blockIdx.y * 32;
  for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
    const int64_t i = i0 + threadIdx.x, j = j0 + jj;
    tile[jj][threadIdx.x] = (i < m && j < n) ? src[j * lda + i] : 0.0;
  }
  __syncthreads();
  for (int ii = threadIdx.y; ii < 32; ii += blockDim.y) {
    const int64_t i = i0 + ii, j = j0 + threadIdx.x;
    if (i < m && j < ld) dst[i * ld + j] = tile[threadIdx.x][ii];
  }
}

__global__ void k_axpby(int64_t n, double a, const double* x, double b, double* y) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
    y[e] = a * x[e] + b * y[e];
}

}  // namespace adaprox
