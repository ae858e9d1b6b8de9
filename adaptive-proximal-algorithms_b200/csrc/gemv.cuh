// gemv.cuh -- the two matrix passes of an iteration as grid-wide device phases.
//
//   gemv_n_phase : zpart[c][i]  = sum_{j in chunk c} A[i][j] x[j]       (A*x,  lasso/runme.jl:22)
//   gemv_t_phase : gpart[b][j]  = sum_{i in CTA b's rows} A[i][j] r[i]  (A'*r, lasso/runme.jl:23)
//
// Both stream the row-major matrix exactly once with 128-bit non-allocating
// loads, 16 of them in flight per thread.  Work units are (column chunk of
// 2048, block of `rb` rows), dealt out contiguously (chunk-major) to the CTAs
// of the persistent grid, so every CTA streams one long contiguous-by-rows
// slab and x / the accumulators stay on chip for the whole slab:
//   A*x : the x chunk lives in shared memory (16 KB), one warp owns one row
//         of the unit at a time -> one shuffle reduction per row, no barrier.
//   A'r : each thread owns 8 columns of the chunk in registers for the whole
//         slab -> no reduction at all; r[i] is a warp-uniform load.
// Partials are combined in a fixed order by the finalize helpers below
// (no atomics: results are reproducible run to run).
#pragma once
#include "common.cuh"
#include "gemv_ring.cuh"

namespace adaprox {

// ---------------------------------------------------------------------------
// dense A*x partials
// ---------------------------------------------------------------------------
__device__ __forceinline__ void load_x_chunk(const DMat& M, const double* x, int c, double* s_x) {
  const int64_t col0 = (int64_t)c * kChunk;
  for (int k = threadIdx.x; k < kChunk; k += kThreads) {
    const int64_t j = col0 + k;
    s_x[k] = (j < M.n) ? ldcg(x + j) : 0.0;
  }
}

__device__ __forceinline__ double row_chunk_dot(const double* __restrict__ arow, const double* s_x, int nvec, int lane) {
  // arow: start of this row's chunk; nvec: number of valid double2 in the chunk
  double acc0 = 0.0, acc1 = 0.0;
  constexpr int kU = 16;
  int k = 0;
  for (; k + kU * 32 <= nvec; k += kU * 32) {
    double2 a[kU];
    ld_stream16_512(arow + 2 * (k + lane), a);          // element u: +u*32 double2 = +u*512 B
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const double2 xv = *reinterpret_cast<const double2*>(s_x + 2 * (k + u * 32 + lane));
      acc0 = fma(a[u].x, xv.x, acc0);
      acc1 = fma(a[u].y, xv.y, acc1);
    }
  }
  if (k < nvec) {   // ragged tail of the last chunk
    double2 a[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int idx = k + u * 32 + lane;
      a[u] = (idx < nvec) ? ld_stream(arow + 2 * idx) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int idx = k + u * 32 + lane;
      if (idx < nvec) {
        const double2 xv = *reinterpret_cast<const double2*>(s_x + 2 * idx);
        acc0 = fma(a[u].x, xv.x, acc0);
        acc1 = fma(a[u].y, xv.y, acc1);
      }
    }
  }
  return warp_sum(acc0 + acc1);
}

__device__ __noinline__ void gemv_n_dense(const DMat& M, const double* x, double* s_x, int b, int G) {
  const int64_t U = (int64_t)M.nchunks * M.nrb;
  const int64_t u0 = unit_begin(U, b, G), u1 = unit_begin(U, b + 1, G);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int cur_c = -1;
  for (int64_t u = u0; u < u1; ++u) {
    const int c = (int)(u / M.nrb);
    const int64_t rbi = u - (int64_t)c * M.nrb;
    if (c != cur_c) {
      __syncthreads();
      load_x_chunk(M, x, c, s_x);
      __syncthreads();
      cur_c = c;
    }
    const int64_t col0 = (int64_t)c * kChunk;
    const int64_t width = (M.ld - col0 < kChunk) ? (M.ld - col0) : kChunk;   // ld is padded -> even
    const int nvec = (int)(width >> 1);
    const int64_t r0 = rbi * M.rb;
    const int64_t r1 = (r0 + M.rb < M.m) ? r0 + M.rb : M.m;
    for (int64_t row = r0 + warp; row < r1; row += kWarps) {
      const double s = row_chunk_dot(M.a + row * M.ld + col0, s_x, nvec, lane);
      if (lane == 0) M.zpart[(int64_t)c * M.m + row] = s;
    }
  }
}

// ---------------------------------------------------------------------------
// dense A'r partials
// ---------------------------------------------------------------------------
__device__ __noinline__ void gemv_t_dense(const DMat& M, const double* r, int b, int G) {
  const int64_t U = (int64_t)M.nchunks * M.nrb;
  const int64_t u0 = unit_begin(U, b, G), u1 = unit_begin(U, b + 1, G);
  constexpr int kH = kV / 2;    // double2 per thread per row
  double2 acc[kH];
  int cur_c = -1;
  double* gout = M.gpart + (int64_t)b * M.npad;

  auto flush = [&](int c) {
    const int64_t col0 = (int64_t)c * kChunk;
#pragma unroll
    for (int k = 0; k < kH; ++k) {
      const int64_t col = col0 + 2 * (k * kThreads + threadIdx.x);
      if (col < M.ld) *reinterpret_cast<double2*>(gout + col) = acc[k];
    }
  };

  for (int64_t u = u0; u < u1; ++u) {
    const int c = (int)(u / M.nrb);
    const int64_t rbi = u - (int64_t)c * M.nrb;
    if (c != cur_c) {
      if (cur_c >= 0) flush(cur_c);
#pragma unroll
      for (int k = 0; k < kH; ++k) acc[k] = make_double2(0.0, 0.0);
      cur_c = c;
    }
    const int64_t col0 = (int64_t)c * kChunk;
    const int64_t r0 = rbi * M.rb;
    const int64_t r1 = (r0 + M.rb < M.m) ? r0 + M.rb : M.m;
    const double* base = M.a + col0 + 2 * threadIdx.x;
    bool ok[kH];
#pragma unroll
    for (int k = 0; k < kH; ++k) ok[k] = (col0 + 2 * (k * kThreads + threadIdx.x)) < M.ld;
    int64_t row = r0;
    if (ok[kH - 1]) {            // full-width chunk: all 16 loads of a 4-row batch in one block
      for (; row + kRowBatch <= r1; row += kRowBatch) {
        double2 a[kRowBatch][kH];
        double rv[kRowBatch];
        const double* p = base + row * M.ld;
        ld_stream4x4_4096(p, p + M.ld, p + 2 * M.ld, p + 3 * M.ld, a);
#pragma unroll
        for (int q = 0; q < kRowBatch; ++q) rv[q] = ldcg(r + row + q);
#pragma unroll
        for (int q = 0; q < kRowBatch; ++q)
#pragma unroll
          for (int k = 0; k < kH; ++k) {
            acc[k].x = fma(a[q][k].x, rv[q], acc[k].x);
            acc[k].y = fma(a[q][k].y, rv[q], acc[k].y);
          }
      }
    }
    for (; row + kRowBatch <= r1; row += kRowBatch) {   // ragged last chunk: predicated loads
      double2 a[kRowBatch][kH];
      double rv[kRowBatch];
#pragma unroll
      for (int q = 0; q < kRowBatch; ++q) {
        const double* p = base + (row + q) * M.ld;
#pragma unroll
        for (int k = 0; k < kH; ++k) a[q][k] = ok[k] ? ld_stream(p + 2 * k * kThreads) : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int q = 0; q < kRowBatch; ++q) rv[q] = ldcg(r + row + q);
#pragma unroll
      for (int q = 0; q < kRowBatch; ++q)
#pragma unroll
        for (int k = 0; k < kH; ++k) {
          acc[k].x = fma(a[q][k].x, rv[q], acc[k].x);
          acc[k].y = fma(a[q][k].y, rv[q], acc[k].y);
        }
    }
    for (; row < r1; ++row) {       // row remainder
      const double* p = base + row * M.ld;
      const double rv = ldcg(r + row);
      double2 a[kH];
#pragma unroll
      for (int k = 0; k < kH; ++k) a[k] = ok[k] ? ld_stream(p + 2 * k * kThreads) : make_double2(0.0, 0.0);
#pragma unroll
      for (int k = 0; k < kH; ++k) {
        acc[k].x = fma(a[k].x, rv, acc[k].x);
        acc[k].y = fma(a[k].y, rv, acc[k].y);
      }
    }
  }
  if (cur_c >= 0) flush(cur_c);
}

// ---------------------------------------------------------------------------
// CSR: LPR lanes per row (32 / LPR rows per warp side by side), 4 nonzeros per lane in flight.
// A row's value is a chain rowptr -> colind -> x[colind] of L2 round trips; with ~30-80 nonzeros per row (rcv1 shape)
// a full warp per row leaves most lanes idle and serialises the rows of a warp, so LPR is chosen per matrix from the
// mean row length (csr_lanes_per_row) and each lane issues its 4 index/value loads, then its 4 gathers, together.
// At 16 lanes per row both rcv1-shaped sweeps move ~2.1 M L2 sectors (0.57 M of CSR data + 1.5 M gathers) in 10.6 / 12.6 us
// = 6.3 / 5.3 TB/s of L2 sector traffic: bound by the gathers' sector granularity.
// The summation order (lane, then position, then the xor butterfly) is fixed: reruns are bit-identical.
// ---------------------------------------------------------------------------
template <int LPR>
__device__ __forceinline__ void spmv_rows_t(int64_t nrows, const int64_t* __restrict__ rowptr, const int* __restrict__ colind,
                                            const double* __restrict__ vals, const double* x, double* out, int b, int G) {
  constexpr int RPW = 32 / LPR;                     // rows per warp side by side
  constexpr int RU = 1;                             // row groups per warp in flight; 2 measured no faster (11.4 / 13.8 vs 10.6 / 12.6 us):
                                                    // the sweeps sit at the L2 sector rate (one 32 B sector per gathered double), not on latency
  const int lane = threadIdx.x & 31, sub = lane / LPR, sl = lane % LPR;
  const int64_t gw = (int64_t)b * kWarps + (threadIdx.x >> 5);
  const int64_t nw = (int64_t)G * kWarps;
  for (int64_t base = gw * RPW; base < nrows; base += nw * RPW * RU) {     // warp-uniform trip count
    int64_t row[RU], k[RU], k1[RU];
    double s[RU];
#pragma unroll
    for (int r = 0; r < RU; ++r) {
      row[r] = base + r * nw * RPW + sub;
      const bool valid = row[r] < nrows;
      k[r] = valid ? rowptr[row[r]] + sl : 0;
      k1[r] = valid ? rowptr[row[r] + 1] : 0;
      s[r] = 0.0;
    }
    for (;;) {
      bool any = false;
#pragma unroll
      for (int r = 0; r < RU; ++r) any = any || (k[r] < k1[r]);
      if (!any) break;
      int c[RU][4]; double v[RU][4], xv[RU][4];
#pragma unroll
      for (int r = 0; r < RU; ++r)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int64_t kk = k[r] + u * LPR;
          const bool ok = kk < k1[r];
          c[r][u] = ok ? colind[kk] : 0;
          v[r][u] = ok ? vals[kk] : 0.0;
        }
#pragma unroll
      for (int r = 0; r < RU; ++r)
#pragma unroll
        for (int u = 0; u < 4; ++u) xv[r][u] = (k[r] + u * LPR < k1[r]) ? ldcg(x + c[r][u]) : 0.0;
#pragma unroll
      for (int r = 0; r < RU; ++r) {
#pragma unroll
        for (int u = 0; u < 4; ++u) s[r] = fma(v[r][u], xv[r][u], s[r]);
        k[r] += 4 * LPR;
      }
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < RU; ++r) {
      double t = s[r];
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (row[r] < nrows && sl == 0) out[row[r]] = t;
    }
  }
}
// measured on the rcv1 shape (tools/csr_sweep.py, profiles/r01_csr_sweep.md): 16 lanes per row are the fastest for both
// 75 nonzeros per row (X*w: 11.6 / 10.6 / 15.3 / 18.0 us with 32 / 16 / 8 / 4 lanes) and 32 per row (X'*r: 26.8 / 12.6 / 19.1 / 16.0)
__host__ __device__ inline int csr_lanes_per_row(int64_t nnz, int64_t nrows) {
  const int64_t avg = nnz / (nrows > 0 ? nrows : 1);
  return avg > 192 ? 32 : (avg > 24 ? 16 : (avg > 12 ? 8 : 4));
}
__device__ __forceinline__ void spmv_rows(int lpr, int64_t nrows, const int64_t* __restrict__ rowptr, const int* __restrict__ colind,
                                          const double* __restrict__ vals, const double* x, double* out, int b, int G) {
  switch (lpr) {
    case 4: spmv_rows_t<4>(nrows, rowptr, colind, vals, x, out, b, G); break;
    case 8: spmv_rows_t<8>(nrows, rowptr, colind, vals, x, out, b, G); break;
    case 16: spmv_rows_t<16>(nrows, rowptr, colind, vals, x, out, b, G); break;
    default: spmv_rows_t<32>(nrows, rowptr, colind, vals, x, out, b, G); break;
  }
}

// ---------------------------------------------------------------------------
// phases
// ---------------------------------------------------------------------------
// Sweep direction (dense ring path).  Every CTA streams its own contiguous slab; when a sweep ends, the L2 holds the part of each
// slab read LAST.  A*x therefore runs in the direction opposite to the previous sweep over the same matrix (its rows are
// independent, so the bits do not change), A'r always runs first row -> last row (its sums over the rows keep their order):
// least squares (A*x, A'r per iteration), the Gram form (Z'x, Z*u) and a dense Quadratic (Q*x only) all alternate, and each sweep
// starts on the tens of MB the previous one left in the 126 MB L2 instead of evicting them before it gets there.
__device__ __forceinline__ void gemv_n_phase(const DMat& M, const double* x, Sh& sh, int b, int G) {
  if (M.kind == MAT_DENSE) {
    if (M.path == 1) {
      const bool rev = M.alternate && ((sh.fwd_last >> M.slot) & 1u);
      gemv_n_ring(M, x, sh, b, G, rev);
      sh.fwd_last = rev ? (sh.fwd_last & ~(1u << M.slot)) : (sh.fwd_last | (1u << M.slot));
    } else gemv_n_dense(M, x, sh.x, b, G);
  } else if (M.kind == MAT_CSR) spmv_rows(M.lpr_n, M.m, M.rowptr, M.colind, M.vals, x, M.zpart, b, G);
}
__device__ __forceinline__ void gemv_t_phase(const DMat& M, const double* r, Sh& sh, int b, int G) {
  if (M.kind == MAT_DENSE) {
    if (M.path == 1) { gemv_t_ring(M, r, sh, b, G); sh.fwd_last |= 1u << M.slot; }
    else gemv_t_dense(M, r, b, G);
  } else if (M.kind == MAT_CSR) spmv_rows(M.lpr_t, M.n, M.t_rowptr, M.t_colind, M.t_vals, r, M.gpart, b, G);
}

// (A*x)_i from the partials (fixed chunk order)
__device__ __forceinline__ double zsum(const DMat& M, int64_t i) {
  if (M.kind == MAT_CSR) return ldcg(M.zpart + i);
  double s = 0.0;
  for (int c0 = 0; c0 < M.nchunks; c0 += 8) {             // loads eight at a time, additions in chunk order
    double v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = c0 + q < M.nchunks ? ldcg(M.zpart + (int64_t)(c0 + q) * M.m + i) : 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) if (c0 + q < M.nchunks) s += v[q];
  }
  return s;
}

// (A'r)_j for j in [j0, j1) from the per-CTA partials (fixed CTA order), written
// to out[j].  Cooperative over the calling CTA; ends with a __syncthreads so the
// caller may read out[j0..j1) back.
__device__ __forceinline__ void gsum_slice(const DMat& M, int64_t j0, int64_t j1, double* out, int G) {
  if (M.kind == MAT_CSR) {
    for (int64_t j = j0 + threadIdx.x; j < j1; j += kThreads) out[j] = ldcg(M.gpart + j);
    __syncthreads();
    return;
  }
  const int64_t U = (int64_t)M.nchunks * M.nrb;
  // a warp per column pays off when MANY CTAs hold a partial of the column (tall matrices: up to nrb owners per chunk); with at
  // most 16 owners a thread per column is faster (1 x N maps of the dual SVM: 37.6 -> ~9 us per A'y at N = 20000)
  const bool warp_per_col = (G > 16 * M.nchunks) && (M.nrb > 16);
  if (!warp_per_col) {
    int cur_c = -1, blo = 0, bhi = -1;
    for (int64_t j = j0 + threadIdx.x; j < j1; j += kThreads) {
      const int c = (int)(j / kChunk);
      if (c != cur_c) {
        blo = unit_owner(U, (int64_t)c * M.nrb, G);
        bhi = unit_owner(U, (int64_t)(c + 1) * M.nrb - 1, G);
        cur_c = c;
      }
      double s = 0.0;
      for (int b0 = blo; b0 <= bhi; b0 += 8) {            // loads eight at a time, additions in CTA order (common.cuh: lane_strided_sum)
        double v[8];
        bool has[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int bb = b0 + q;
          has[q] = bb <= bhi && unit_begin(U, bb + 1, G) > unit_begin(U, bb, G);
          v[q] = has[q] ? ldcg(M.gpart + (int64_t)bb * M.npad + j) : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) if (has[q]) s += v[q];
      }
      out[j] = s;
    }
  } else {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t j = j0 + warp; j < j1; j += kWarps) {
      const int c = (int)(j / kChunk);
      const int blo = unit_owner(U, (int64_t)c * M.nrb, G);
      const int bhi = unit_owner(U, (int64_t)(c + 1) * M.nrb - 1, G);
      double s = 0.0;
      for (int b0 = blo + lane; b0 <= bhi; b0 += 256) {
        double v[8];
        bool has[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int bb = b0 + 32 * q;
          has[q] = bb <= bhi && unit_begin(U, bb + 1, G) > unit_begin(U, bb, G);
          v[q] = has[q] ? ldcg(M.gpart + (int64_t)bb * M.npad + j) : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) if (has[q]) s += v[q];
      }
      s = warp_sum(s);
      if (lane == 0) out[j] = s;
    }
  }
  __syncthreads();
}

}  // namespace adaprox
