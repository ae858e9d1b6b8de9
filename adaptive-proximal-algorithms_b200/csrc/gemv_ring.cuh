// gemv_ring.cuh -- the matrix passes with the stream staged through shared memory
// by the bulk-copy engine (cp.async.bulk + mbarrier, SASS: UBLKCP / SYNCS).
//
// Why: with register-staged loads the number of bytes in flight per SM is tied to
// the registers ptxas is willing to spend; inside the big persistent kernel it
// serialises the loads (4-6 x LDG.128 in flight per thread) and the passes lose
// ~20 % of the HBM bandwidth (profiles/r01_notes.md).  Here one elected thread
// keeps kStages-1 row tiles (16 KB each) in flight per CTA, independent of the
// compiler; the 256 threads consume a tile with conflict-free LDS.128 and keep
// x (for A*x) or the column accumulators (for A'r) in registers.
//
// A tile is one row of one 2048-column chunk (<= 16 KB, contiguous in the
// row-major layout, 128-byte aligned).  Thread t owns columns 2*(k*256 + t),
// +1 for k = 0..3 of the chunk in BOTH passes.
#pragma once
#include "common.cuh"

namespace adaprox {

constexpr int kStages = 4;
constexpr int kStageBytes = kChunk * 8;                 // 16 KB
constexpr int kRingBytes = kStages * kStageBytes;       // 64 KB of dynamic shared memory per CTA (2 CTAs/SM leave ~90 KB of L1)
constexpr int kPartRows = 64;                           // rows between cross-warp combines in A*x

struct Sh {
  double* x;            // LDG path: x chunk (aliases stage 0 of the ring)
  double* scr;          // reduction scratch
  double* part;         // [kPartRows][kWarps] per-warp row partials of A*x
  unsigned char* ring_gen;
  uint32_t ring;        // shared-space address of stage 0
  uint32_t full;        // shared-space address of full[0]  (8 B apart)
  uint32_t empty;       // shared-space address of empty[0]
  uint32_t count;       // tiles this CTA has pushed through the ring so far (uniform across the CTA)
  uint32_t fwd_last;    // bit DMat::slot: the last dense sweep over that matrix ran first row -> last row (see gemv_n_phase)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t addr, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  do {
#ifdef ADAPROX_RING_POLL
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
#else
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
#endif
  } while (!ok);
}
// global -> shared bulk copy completing on an mbarrier (bytes: multiple of 16; both addresses 16-byte aligned)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}

// the same with an L2 eviction policy (createpolicy): streaming data that is read once should not displace lines somebody re-reads
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar), "l"(policy) : "memory");
}

// once per kernel, by all threads
__device__ __forceinline__ void sh_init(Sh& sh, unsigned char* dyn, double* scr, double* part, unsigned long long* bars) {
  sh.x = reinterpret_cast<double*>(dyn);
  sh.scr = scr;
  sh.part = part;
  sh.ring_gen = dyn;
  sh.ring = smem_u32(dyn);
  sh.full = smem_u32(bars);
  sh.empty = smem_u32(bars + kStages);
  sh.count = 0;
  sh.fwd_last = 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(sh.full + 8 * s, 1); mbar_init(sh.empty + 8 * s, kWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
}

// explicit shared-space accesses (a pointer that travelled through a struct would compile to generic LD/ST)
__device__ __forceinline__ double2 lds2(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts1(uint32_t addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }
__device__ __forceinline__ double lds1(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}

// The tiles of CTA b: its contiguous (chunk-major) unit range is a first segment (chunk c0, rows
// [row0, end0)), whole chunks c0+1 .. c1-1 (rows [0, m)) and a last segment (chunk c1, rows [0, end1)).
struct Segs {
  int c0, c1;
  int64_t row0, end0, end1, m, T;
  __device__ __forceinline__ void init(int nchunks, int64_t nrb, int rb, int64_t mrows, int b, int G) {
    const int64_t U = (int64_t)nchunks * nrb;
    const int64_t u0 = unit_begin(U, b, G), u1 = unit_begin(U, b + 1, G);
    m = mrows; T = 0; c0 = 0; c1 = -1; row0 = end0 = end1 = 0;
    if (u0 >= u1) return;
    c0 = (int)(u0 / nrb);
    c1 = (int)((u1 - 1) / nrb);
    row0 = (u0 - (int64_t)c0 * nrb) * rb;
    int64_t le = ((u1 - 1) - (int64_t)c1 * nrb + 1) * rb;
    end1 = le < m ? le : m;
    end0 = (c0 == c1) ? end1 : m;
    T = (end0 - row0) + ((c1 > c0) ? (int64_t)(c1 - c0 - 1) * m + end1 : 0);
  }
  __device__ __forceinline__ int64_t seg_begin(int c) const { return c == c0 ? row0 : 0; }
  __device__ __forceinline__ int64_t seg_end(int c) const { return c == c0 ? end0 : (c == c1 ? end1 : m); }
};

// producer cursor (meaningful in thread 0 only)
struct Prod {
  int c;
  int64_t row, end;
  uint32_t g;            // running ring index of the next tile to issue
  int64_t left;          // tiles still to issue
  int64_t keep;          // the last `keep` tiles of the sweep are loaded evict_last, the others evict_first; < 0: no hints
  uint64_t pol_first, pol_last;
  __device__ __forceinline__ void policies(int64_t keep_tiles) {
    keep = keep_tiles;
    if (keep >= 0) { pol_first = l2_policy_evict_first(); pol_last = l2_policy_evict_last(); }
  }
};

// rev: the slab is walked last tile -> first tile; p.end is then the first row of the current segment (inclusive)
__device__ __forceinline__ void prod_issue(Prod& p, const Segs& sg, const double* a, int64_t ld, uint32_t ring, uint32_t full,
                                           uint32_t empty, bool rev = false) {
  const uint32_t s = p.g % kStages, ph = (p.g / kStages) & 1u;
  mbar_wait(empty + 8 * s, ph ^ 1u);
  const int64_t w = ld - (int64_t)p.c * kChunk;
  const uint32_t bytes = (uint32_t)((w < kChunk ? w : kChunk) * 8);
  mbar_expect_tx(full + 8 * s, bytes);
  // What the NEXT sweep over this matrix reads first is what this one reads last (gemv_n_phase): those tiles are asked to stay in
  // L2, everything else (streamed once, or just re-read from L2) to leave it first.
  if (p.keep < 0) bulk_g2s(ring + s * kStageBytes, a + p.row * ld + (int64_t)p.c * kChunk, bytes, full + 8 * s);
  else bulk_g2s_hint(ring + s * kStageBytes, a + p.row * ld + (int64_t)p.c * kChunk, bytes, full + 8 * s,
                     p.left <= p.keep ? p.pol_last : p.pol_first);
  ++p.g; --p.left;
  if (!rev) {
    if (++p.row >= p.end) { ++p.c; p.row = 0; p.end = sg.seg_end(p.c); }
  } else if (--p.row < p.end) {
    --p.c; p.row = sg.seg_end(p.c) - 1; p.end = sg.seg_begin(p.c);
  }
}

// ---------------------------------------------------------------------------
// A*x partials: zpart[c][row] = sum over the chunk
// rev: the CTA walks its slab from the last tile to the first.  A row's value does not depend on the order the rows are
// visited in (one fixed-order sum of kWarps warp partials per row), so both directions leave the same bits; what changes is
// which end of the slab is still in L2 from the previous sweep over the same matrix (gemv_n_phase picks the direction).
// ---------------------------------------------------------------------------
__device__ __noinline__ void gemv_n_ring(const DMat& M, const double* x, Sh& sh, int b, int G, bool rev) {
  constexpr int kH = kV / 2;
  // descriptors into registers (nothing below may alias them)
  const double* const a = M.a;
  const int64_t ld = M.ld, n = M.n, m = M.m;
  double* const zpart = M.zpart;
  const uint32_t ring = sh.ring, full = sh.full, empty = sh.empty, part = smem_u32(sh.part);
  Segs sg;
  sg.init(M.nchunks, M.nrb, M.rb, m, b, G);
  if (sg.T == 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t g = sh.count;                  // ring index of the next tile to consume
  Prod p;
  p.g = g; p.left = sg.T; p.policies(M.keep);
  if (!rev) { p.c = sg.c0; p.row = sg.row0; p.end = sg.end0; }
  else { p.c = sg.c1; p.row = sg.seg_end(sg.c1) - 1; p.end = sg.seg_begin(sg.c1); }
  if (threadIdx.x == 0)
    for (int k = 0; k < kStages - 1 && p.left > 0; ++k) prod_issue(p, sg, a, ld, ring, full, empty, rev);

  const int nc = sg.c1 - sg.c0 + 1;
  for (int ci = 0; ci < nc; ++ci) {
    const int c = rev ? sg.c1 - ci : sg.c0 + ci;
    const int64_t col0 = (int64_t)c * kChunk;
    double2 xr[kH];
    bool ok[kH];
#pragma unroll
    for (int k = 0; k < kH; ++k) {
      const int64_t j = col0 + 2 * (k * kThreads + threadIdx.x);
      ok[k] = j < ld;
      xr[k].x = (j < n) ? ldcg(x + j) : 0.0;
      xr[k].y = (j + 1 < n) ? ldcg(x + j + 1) : 0.0;
    }
    const int64_t rbeg = sg.seg_begin(c), rend = sg.seg_end(c);
    for (int64_t done = 0; done < rend - rbeg; done += kPartRows) {   // tile q of the block: row rbeg+done+q, or rend-1-done-q
      const int nrows = (int)((rend - rbeg - done < kPartRows) ? (rend - rbeg - done) : kPartRows);
      for (int q = 0; q < nrows; ++q) {
        if (threadIdx.x == 0 && p.left > 0) prod_issue(p, sg, a, ld, ring, full, empty, rev);
        const uint32_t s = g % kStages, ph = (g / kStages) & 1u;
        mbar_wait(full + 8 * s, ph);
        const uint32_t tile = ring + s * kStageBytes + threadIdx.x * 16;
        double p0 = 0.0, p1 = 0.0;
#pragma unroll
        for (int k = 0; k < kH; ++k)
          if (ok[k]) {
            const double2 av = lds2(tile + k * kThreads * 16);
            p0 = fma(av.x, xr[k].x, p0);
            p1 = fma(av.y, xr[k].y, p1);
          }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + 8 * s);
        const double ps = warp_sum(p0 + p1);
        if (lane == 0) sts1(part + (q * kWarps + warp) * 8, ps);
        ++g;
      }
      __syncthreads();                     // combine the per-warp partials of this block of rows, fixed warp order
      if (threadIdx.x < nrows) {
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) sum += lds1(part + (threadIdx.x * kWarps + w) * 8);
        const int64_t row = rev ? rend - 1 - done - threadIdx.x : rbeg + done + threadIdx.x;
        zpart[(int64_t)c * m + row] = sum;
      }
      __syncthreads();
    }
  }
  sh.count = g;
}

// ---------------------------------------------------------------------------
// A'r partials: gpart[b][col] = sum over this CTA's rows
// ---------------------------------------------------------------------------
__device__ __noinline__ void gemv_t_ring(const DMat& M, const double* r, Sh& sh, int b, int G) {
  constexpr int kH = kV / 2;
  const double* const a = M.a;
  const int64_t ld = M.ld, m = M.m;
  double* const gout = M.gpart + (int64_t)b * M.npad;
  const uint32_t ring = sh.ring, full = sh.full, empty = sh.empty;
  Segs sg;
  sg.init(M.nchunks, M.nrb, M.rb, m, b, G);
  if (sg.T == 0) return;
  const int lane = threadIdx.x & 31;
  uint32_t g = sh.count;
  Prod p;
  p.c = sg.c0; p.row = sg.row0; p.end = sg.end0; p.g = g; p.left = sg.T; p.policies(M.keep);
  if (threadIdx.x == 0)
    for (int k = 0; k < kStages - 1 && p.left > 0; ++k) prod_issue(p, sg, a, ld, ring, full, empty);

  for (int c = sg.c0; c <= sg.c1; ++c) {
    const int64_t col0 = (int64_t)c * kChunk;
    double2 acc[kH];
    bool ok[kH];
#pragma unroll
    for (int k = 0; k < kH; ++k) { acc[k] = make_double2(0.0, 0.0); ok[k] = (col0 + 2 * (k * kThreads + threadIdx.x)) < ld; }
    const int64_t rbeg = sg.seg_begin(c), rend = sg.seg_end(c);
    for (int64_t blk = rbeg; blk < rend; blk += 32) {
      const int nrows = (int)((rend - blk < 32) ? (rend - blk) : 32);
      const double rblk = (lane < nrows) ? ldcg(r + blk + lane) : 0.0;      // lane l holds r[blk + l]
      for (int q = 0; q < nrows; ++q) {
        if (threadIdx.x == 0 && p.left > 0) prod_issue(p, sg, a, ld, ring, full, empty);
        const double rv = __shfl_sync(0xffffffffu, rblk, q);
        const uint32_t s = g % kStages, ph = (g / kStages) & 1u;
        mbar_wait(full + 8 * s, ph);
        const uint32_t tile = ring + s * kStageBytes + threadIdx.x * 16;
#pragma unroll
        for (int k = 0; k < kH; ++k)
          if (ok[k]) {
            const double2 av = lds2(tile + k * kThreads * 16);
            acc[k].x = fma(av.x, rv, acc[k].x);
            acc[k].y = fma(av.y, rv, acc[k].y);
          }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + 8 * s);
        ++g;
      }
    }
#pragma unroll
    for (int k = 0; k < kH; ++k) {
      const int64_t col = col0 + 2 * (k * kThreads + threadIdx.x);
      if (col < ld) *reinterpret_cast<double2*>(gout + col) = acc[k];
    }
  }
  sh.count = g;
}

}  // namespace adaprox
