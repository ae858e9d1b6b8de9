// common.cuh -- device-side descriptors and helpers shared by every kernel.
//
// Layout in HBM:
//   dense matrices are ROW-major with the leading dimension padded to a
//   multiple of 16 doubles (128 B) and zero-filled padding, so every row
//   starts on a 128-byte line and 16-byte vector loads never straddle a row;
//   CSR matrices keep both CSR(X) and CSR(X') so X'v is a gather, not a
//   scatter (no fp64 atomics anywhere: all reductions are fixed-order).
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include <math.h>

namespace cg = cooperative_groups;

namespace adaprox {

constexpr int kThreads = 256;           // threads per CTA of every solver kernel
constexpr int kWarps = kThreads / 32;
constexpr int kV = 8;                    // matrix elements per thread per row of a chunk
constexpr int kChunk = kThreads * kV;    // 2048 columns per column chunk (16 KB per row)
constexpr int kRowBatch = 4;             // rows in flight per thread (16 x LDG.128)
constexpr int kMaxRed = 16;              // reduction slots

enum { MAT_NONE = 0, MAT_DENSE = 1, MAT_CSR = 2 };

struct DMat {
  int kind;
  int64_t m, n, ld;          // local rows, columns, leading dimension (dense)
  const double* a;           // dense row-major [m][ld]
  // CSR(X) and CSR(X')
  const int64_t* rowptr; const int* colind; const double* vals;
  const int64_t* t_rowptr; const int* t_colind; const double* t_vals;
  int64_t nnz;
  int lpr_n, lpr_t;          // CSR: lanes per row of the X*w and X'*r sweeps (csr_lanes_per_row)
  // dense work partition: units = (column chunk c, row block rb), chunk-major
  int nchunks;               // ceil(n / kChunk)
  int rb;                    // rows per unit (multiple of kRowBatch)
  int64_t nrb;               // ceil(m / rb)
  // partial-result buffers owned by the matrix
  double* zpart;             // [nchunks][m]    A*x partials
  double* gpart;             // [grid][npad]    A'r partials, one row per CTA
  int64_t npad;              // n rounded up to kChunk
  int path;                  // dense passes: 1 = bulk-copy shared-memory ring (default), 0 = register-staged LDG.128
  int64_t keep;              // ring path: tiles per CTA a sweep leaves in L2 for the next one (evict_last), < 0: no eviction hints
  int alternate;             // ring path: A*x runs against the direction of the previous sweep (0: ADAPROX_SWEEP_ONE_WAY=1, A/B and tests)
  int slot;                  // 0 = the matrix of f, 1 = the linear map A (index of the sweep-direction bit in Sh::fwd_last)
};

struct DProx {
  int kind, conjugate;
  double lambda, lo, hi;
  const double* lo_vec; const double* hi_vec; const double* shift;
};

// ---------------------------------------------------------------------------
// Julia scalar semantics
// ---------------------------------------------------------------------------
__host__ __device__ inline double jl_min(double a, double b) {
  // Julia's min propagates NaN; C fmin does not (src/AdaProx.jl:228,263,303,517)
  if (a != a || b != b) return NAN;
  return b < a ? b : a;
}
__host__ __device__ inline double jl_max(double a, double b) {
  if (a != a || b != b) return NAN;
  return b > a ? b : a;
}
__host__ __device__ inline double nan_to_zero(double v) { return v != v ? 0.0 : v; }   // src/AdaProx.jl:24
__host__ __device__ inline double sq(double v) { return v * v; }
// norm(v)^2 as Julia evaluates it: a square root followed by a square.
__host__ __device__ inline double norm_sq_jl(double sumsq) { double nrm = sqrt(sumsq); return nrm * nrm; }

// ---------------------------------------------------------------------------
// loads
// ---------------------------------------------------------------------------
// Streaming 128-bit load of matrix data that is read-only for the whole kernel:
// non-coherent path, no L1 allocation (each element is used exactly once).
__device__ __forceinline__ double2 ld_stream(const double* p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

// 16 streaming 128-bit loads issued back to back from ONE asm block, so ptxas
// cannot interleave the consuming FMAs between them (it does when it is short of
// registers, which leaves only 4-6 loads in flight per thread and costs ~20 % of
// the HBM bandwidth -- profiles/r01_notes.md).  `stride_bytes` apart from p.
__device__ __forceinline__ void ld_stream16_512(const double* p, double2 (&a)[16]) {
  asm volatile(
      "ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%32];\n\t"
      "ld.global.nc.L1::no_allocate.v2.f64 {%2, %3}, [%32+512];\n\t"
      "ld.global.nc.L1::no_allocate.v2.f64 {%4, %5}, [%32+1024];\n\t"
      "ld.global.nc.L1::no_allocate.v2.f64 {%6, %7}, [%32+1536];\n\t"
      "ld.global.nc.L1::no_allocate.v2.f64 {%8, %9}, [%32+2048];\n\t"
      "ld.global.nc.L1::no_allocate.v2.f64 {%10, %11}, [%32+2560];\n\t"
      "ld.global.nc.L1::no_allocate.v2.f64 {%12, %13}, [%32+3072];\n\t"
      "ld.global.nc.L1::no_allocate.v2.f64 {%14, %15}, [%32+3584];\n\t"
      "ld.global.nc.L1::no_allocate.v2.f64 {%16, %17}, [%32+4096];\n\t"
      "ld.global.nc.L1::no_allocate.v2.f64 {%18, %19}, [%32+4608];\n\t"
      "ld.global.nc.L1::no_allocate.v2.f64 {%20, %21}, [%32+5120];\n\t"
      "ld.global.nc.L1::no_allocate.v2.f64 {%22, %23}, [%32+5632];\n\t"
      "ld.global.nc.L1::no_allocate.v2.f64 {%24, %25}, [%32+6144];\n\t"
      "ld.global.nc.L1::no_allocate.v2.f64 {%26, %27}, [%32+6656];\n\t"
      "ld.global.nc.L1::no_allocate.v2.f64 {%28, %29}, [%32+7168];\n\t"
      "ld.global.nc.L1::no_allocate.v2.f64 {%30, %31}, [%32+7680];"
      : "=d"(a[0].x), "=d"(a[0].y), "=d"(a[1].x), "=d"(a[1].y), "=d"(a[2].x), "=d"(a[2].y), "=d"(a[3].x), "=d"(a[3].y),
        "=d"(a[4].x), "=d"(a[4].y), "=d"(a[5].x), "=d"(a[5].y), "=d"(a[6].x), "=d"(a[6].y), "=d"(a[7].x), "=d"(a[7].y),
        "=d"(a[8].x), "=d"(a[8].y), "=d"(a[9].x), "=d"(a[9].y), "=d"(a[10].x), "=d"(a[10].y), "=d"(a[11].x), "=d"(a[11].y),
        "=d"(a[12].x), "=d"(a[12].y), "=d"(a[13].x), "=d"(a[13].y), "=d"(a[14].x), "=d"(a[14].y), "=d"(a[15].x), "=d"(a[15].y)
      : "l"(p));
}
// 4 rows x 4 loads (4096 B apart within a row): the A'r batch of one thread.
__device__ __forceinline__ void ld_stream4x4_4096(const double* p0, const double* p1, const double* p2, const double* p3,
                                                  double2 (&a)[4][4]) {
  asm volatile(
      "ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%32];\n\t"
      "ld.global.nc.L1::no_allocate.v2.f64 {%2, %3}, [%32+4096];\n\t"
      "ld.global.nc.L1::no_allocate.v2.f64 {%4, %5}, [%32+8192];\n\t"
      "ld.global.nc.L1::no_allocate.v2.f64 {%6, %7}, [%32+12288];\n\t"
      "ld.global.nc.L1::no_allocate.v2.f64 {%8, %9}, [%33];\n\t"
      "ld.global.nc.L1::no_allocate.v2.f64 {%10, %11}, [%33+4096];\n\t"
      "ld.global.nc.L1::no_allocate.v2.f64 {%12, %13}, [%33+8192];\n\t"
      "ld.global.nc.L1::no_allocate.v2.f64 {%14, %15}, [%33+12288];\n\t"
      "ld.global.nc.L1::no_allocate.v2.f64 {%16, %17}, [%34];\n\t"
      "ld.global.nc.L1::no_allocate.v2.f64 {%18, %19}, [%34+4096];\n\t"
      "ld.global.nc.L1::no_allocate.v2.f64 {%20, %21}, [%34+8192];\n\t"
      "ld.global.nc.L1::no_allocate.v2.f64 {%22, %23}, [%34+12288];\n\t"
      "ld.global.nc.L1::no_allocate.v2.f64 {%24, %25}, [%35];\n\t"
      "ld.global.nc.L1::no_allocate.v2.f64 {%26, %27}, [%35+4096];\n\t"
      "ld.global.nc.L1::no_allocate.v2.f64 {%28, %29}, [%35+8192];\n\t"
      "ld.global.nc.L1::no_allocate.v2.f64 {%30, %31}, [%35+12288];"
      : "=d"(a[0][0].x), "=d"(a[0][0].y), "=d"(a[0][1].x), "=d"(a[0][1].y), "=d"(a[0][2].x), "=d"(a[0][2].y), "=d"(a[0][3].x), "=d"(a[0][3].y),
        "=d"(a[1][0].x), "=d"(a[1][0].y), "=d"(a[1][1].x), "=d"(a[1][1].y), "=d"(a[1][2].x), "=d"(a[1][2].y), "=d"(a[1][3].x), "=d"(a[1][3].y),
        "=d"(a[2][0].x), "=d"(a[2][0].y), "=d"(a[2][1].x), "=d"(a[2][1].y), "=d"(a[2][2].x), "=d"(a[2][2].y), "=d"(a[2][3].x), "=d"(a[2][3].y),
        "=d"(a[3][0].x), "=d"(a[3][0].y), "=d"(a[3][1].x), "=d"(a[3][1].y), "=d"(a[3][2].x), "=d"(a[3][2].y), "=d"(a[3][3].x), "=d"(a[3][3].y)
      : "l"(p0), "l"(p1), "l"(p2), "l"(p3));
}
static_assert(kThreads == 256 && kV == 8 && kRowBatch == 4, "the fused load blocks assume 256 threads x 8 columns, 4-row batches");

// Data produced by other CTAs earlier in the same (persistent) kernel: read
// through L2 (ld.global.cg) so a stale L1 line can never be observed.
__device__ __forceinline__ double ldcg(const double* p) { return __ldcg(p); }
__device__ __forceinline__ double2 ldcg2(const double* p) { return __ldcg(reinterpret_cast<const double2*>(p)); }

// ---------------------------------------------------------------------------
// deterministic reductions
// ---------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;     // xor butterfly: every lane holds the same bits
}

// Block-reduce K values and store them as this CTA's partials:
//   red[(slot0 + k) * G + blockIdx.x]
template <int K>
__device__ __forceinline__ void block_reduce_store(double (&v)[K], double* red, int G, int slot0, double* s_scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double s = warp_sum(v[k]);
    if (lane == 0) s_scratch[warp * K + k] = s;
  }
  __syncthreads();
  if (threadIdx.x < K) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += s_scratch[w * K + threadIdx.x];
    red[(int64_t)(slot0 + threadIdx.x) * G + blockIdx.x] = s;
  }
  __syncthreads();
}

// sum of p[b * stride] over b = lane, lane + 32, ... < count, added in that order.  The loads are issued eight at a time: written
// as `s += ldcg(...)` in a loop, each addition waits for its own L2 round trip (~0.7 us) before the next load is even issued --
// ten of them in a row for the 296 partials of one scalar.  Same order of additions as the plain loop, so the same bits.
__device__ __forceinline__ double lane_strided_sum(const double* p, int count, int64_t stride, int lane) {
  double s = 0.0;
  for (int b0 = lane; b0 < count; b0 += 256) {
    double v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { const int b = b0 + 32 * q; v[q] = b < count ? ldcg(p + (int64_t)b * stride) : 0.0; }
#pragma unroll
    for (int q = 0; q < 8; ++q) if (b0 + 32 * q < count) s += v[q];
  }
  return s;
}

// Sum the per-CTA partials of K slots in a fixed order; every thread of every
// CTA obtains bit-identical totals.  Must follow a grid-wide sync (or a kernel
// boundary) after the block_reduce_store that produced the partials.
template <int K>
__device__ __forceinline__ void grid_totals(const double* red, int G, int slot0, double (&out)[K], double* s_scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = warp; k < K; k += kWarps) {
    const double s = warp_sum(lane_strided_sum(red + (int64_t)(slot0 + k) * G, G, 1, lane));
    if (lane == 0) s_scratch[k] = s;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) out[k] = s_scratch[k];
  __syncthreads();
}

// contiguous slice [lo, hi) of `len` items owned by CTA b of G
__device__ __forceinline__ void cta_slice(int64_t len, int b, int G, int64_t& lo, int64_t& hi) {
  lo = (len * b) / G;
  hi = (len * (b + 1)) / G;
}

// First unit of CTA b when U units are dealt out contiguously to G CTAs.
__host__ __device__ inline int64_t unit_begin(int64_t U, int b, int G) { return (U * (int64_t)b) / G; }
// The CTA that owns unit u.
__host__ __device__ inline int unit_owner(int64_t U, int64_t u, int G) {
  int b = (int)(((u + 1) * (int64_t)G - 1) / U);
  if (b >= G) b = G - 1;
  while (b + 1 < G && unit_begin(U, b + 1, G) <= u) ++b;
  while (b > 0 && unit_begin(U, b, G) > u) --b;
  return b;
}

}  // namespace adaprox
