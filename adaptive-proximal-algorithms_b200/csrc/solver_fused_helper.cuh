// solver_fused_helper.cuh -- helper CTAs for the single-sweep kernel (included by solver_fused.cuh, variant 1).
//
// k_adapgm_fused needs clusters of 16 CTAs at n = 131072 and the hardware hands out 7 of them: 112 of 148 SMs
// (profiles/r02_fp64_peaks.jsonl).  k_adapgm_helper is a plain (cluster-less) launch on a second stream that lands on the
// stranded SMs and works through the SAME chunk dispenser:
//   * 16 helper CTAs form a "virtual cluster"; CTA rho owns the same 8192 columns a cluster rank owns;
//   * a chunk is processed in batches of R rows.  Pass A streams the batch through the bulk-copy ring and the dot warps leave
//     their warp partials -- the identical code and summation tree as the cluster kernel -- in global memory; the 16 CTAs meet
//     at a global-memory barrier; every CTA sums the 16 x 8 partials of each row in the cluster kernel's order (bit-identical
//     r_i); pass B streams the batch AGAIN (from L2: it was read microseconds ago) and the update warps accumulate
//     acc += A[i, cols] r_i.  A helper SM therefore ingests every row twice, but neither pass holds a ring slot across an
//     exchange, and DRAM still sees every row once.
//   * per-chunk outputs (gpartf[chunk], fpart[chunk]) are the same bits whoever produced them, so reruns stay bit-identical.
// Protocol with the main kernel (FusedArgs: next / go / done): the dispenser value is (sweep << 32) | index; a helper accepts
// whatever sweep its grab belongs to, waits until `go` >= that sweep (the iterate xb[sweep % 3] is complete), and counts the
// chunk in `done`, which the main CTAs wait for after their sweep.  The helpers are optional at every moment: if they are not
// resident (another kernel holds the SMs, a profiler serialises the launches) the roll call times out, they exit, and the main
// clusters process every chunk themselves.
#pragma once

namespace adaprox {

#if ADAPROX_FUSED_VARIANT == 1

constexpr unsigned long long kHelperRollCallNs = 50000000ull;        // 50 ms to get all helper CTAs resident
constexpr unsigned long long kHelperIdleNs = 60000000000ull;         // give up after 60 s without a new sweep (main kernel gone)
constexpr unsigned long long kExitTag = 0xffffffffull;

struct VcBar {                       // barrier of the 16 CTAs of a virtual cluster on a global arrival counter
  unsigned long long* ctr;
  unsigned long long target;
  int* err;
  __device__ __forceinline__ void sync() {
    __syncthreads();
    if (threadIdx.x == 0) {
      target += kFMaxCluster;
      __threadfence();
      atomicAdd(ctr, 1ull);
      unsigned long long seen, t0 = 0;
      for (unsigned spin = 0;; ++spin) {
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(ctr) : "memory");
        if (seen >= target) break;
        if ((spin & 4095u) == 4095u) {
          if (*reinterpret_cast<volatile int*>(err)) break;
          const unsigned long long now = globaltimer_ns();
          if (t0 == 0) t0 = now;
          else if (now - t0 > kGridBarTimeoutNs) { *reinterpret_cast<volatile int*>(err) = 1; __threadfence(); break; }
        }
      }
    }
    __syncthreads();
  }
};

__global__ void __launch_bounds__(kFThreads, 1) k_adapgm_helper(DProblem P, DWork W, FusedArgs fa) {
  extern __shared__ __align__(1024) unsigned char dyn_smem[];
  __shared__ __align__(16) unsigned long long s_bars[2 * kFStages];
  __shared__ double s_r[64];                          // r_i of the current batch (hR <= 64)
  __shared__ unsigned long long s_mail;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int vc = blockIdx.x / kFMaxCluster;
  const uint32_t rho = blockIdx.x % kFMaxCluster;
  const int C = fa.C, R = fa.hR;
  const DMat& M = P.F;
  const uint32_t ring = smem_u32(dyn_smem), bars = smem_u32(s_bars);
  const uint32_t full0 = bars, empty0 = bars + 8 * kFStages;
  unsigned long long* roll = fa.hsync;
  unsigned long long* abortf = fa.hsync + 1;
  unsigned long long* mailbox = fa.hsync + 2 + fa.hV + vc;
  VcBar vb{fa.hsync + 2 + vc, 0ull, fa.err};

  // ---- roll call: every helper CTA of the launch must be resident, or nobody helps ---------------------------------------
  __shared__ int s_ok;
  if (t == 0) {
    const unsigned long long want = (unsigned long long)gridDim.x;
    atomicAdd(roll, 1ull);
    const unsigned long long t0 = globaltimer_ns();
    unsigned long long seen;
    int ok = 1;
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(roll) : "memory");
      if (seen >= want) break;
      if (globaltimer_ns() - t0 > kHelperRollCallNs) { atomicExch(abortf, 1ull); ok = 0; break; }
      __nanosleep(200);
    }
    unsigned long long ab;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(ab) : "l"(abortf) : "memory");
    s_ok = (ok && ab == 0ull) ? 1 : 0;
  }
  __syncthreads();
  if (!s_ok) return;

  if (t == 0) {
    for (int s = 0; s < kFStages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, kFGWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    // the last column slice may be ragged: the bulk copies never touch the tail of its ring slots, zero it once
    int64_t width = M.ld - (int64_t)rho * kFCols;
    width = width < 0 ? 0 : (width > kFCols ? kFCols : width);
    for (int s = 0; s < kFStages; ++s)
      for (int64_t j = width + t; j < kFCols; j += kFThreads) sts1(ring + s * kFStageBytes + (uint32_t)j * 8, 0.0);
  }
  __syncthreads();

  const int64_t col0 = (int64_t)rho * kFCols;
  uint32_t bytes;
  {
    int64_t width = M.ld - col0;
    width = width < 0 ? 0 : (width > kFCols ? kFCols : width);
    bytes = (uint32_t)(width * 8);
  }
  const int64_t ldb = M.ld * 8;
  const int nval = C * kFGWarps;
  uint32_t g = 0;                                     // rows pushed through the ring so far, modulo 6 (slot = g % 3, phase = (g / 3) & 1)
  double* xch = fa.hxch + (size_t)vc * 2 * (size_t)R * (kFMaxCluster * kFGWarps);
  unsigned batch_no = 0;
  unsigned long long idle_t0 = 0;

  for (;;) {
    // ---- take a chunk for the whole virtual cluster -------------------------------------------------------------------
    if (rho == 0 && t == 0) {
      // peek first: near the end of a sweep a helper chunk (two passes, ~1.7x a cluster's time) would finish after the clusters
      // and the whole grid would wait for it -- leave the last hHold chunks to the clusters (reported as "handed out")
      unsigned long long v;
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(fa.next) : "memory");
      if ((v >> 32) != kExitTag && (long long)(v & 0xffffffffull) < (long long)fa.nchunks - fa.hHold) v = atomicAdd(fa.next, 1ull);
      else if ((v >> 32) != kExitTag) v = (v & ~0xffffffffull) | (unsigned long long)fa.nchunks;
      asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(mailbox), "l"(v) : "memory");
    }
    vb.sync();
    if (t == 0) { unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(mailbox) : "memory"); s_mail = v; }
    __syncthreads();
    const unsigned long long v = s_mail;
    const unsigned long long tag = v >> 32;
    const long long c = (long long)(v & 0xffffffffull);
    if (tag == kExitTag || *reinterpret_cast<volatile int*>(fa.err)) break;
    if (c >= fa.nchunks) {
      // this sweep is handed out: wait until the dispenser is armed for another one (or carries the exit tag)
      if (t == 0) {
        if (idle_t0 == 0) idle_t0 = globaltimer_ns();
        unsigned long long cur;
        int give_up = 0;
        for (;;) {
          asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(cur) : "l"(fa.next) : "memory");
          if ((cur >> 32) != tag) break;
          if (globaltimer_ns() - idle_t0 > kHelperIdleNs) { give_up = 1; break; }
          __nanosleep(500);
        }
        s_ok = give_up ? 0 : 1;
      }
      __syncthreads();
      if (!s_ok) break;                               // (uniform over the virtual cluster only approximately: the barrier below is bounded)
      continue;
    }
    idle_t0 = 0;
    // ---- the iterate of that sweep must be complete ---------------------------------------------------------------------
    if (t == 0) {
      unsigned long long cur;
      for (unsigned spin = 0;; ++spin) {
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(cur) : "l"(fa.go) : "memory");
        if (cur >= tag) break;
        if ((spin & 1023u) == 1023u && *reinterpret_cast<volatile int*>(fa.err)) break;
      }
    }
    __syncthreads();
    const double* x = W.xb[tag % 3];
    const int64_t r0 = c * (int64_t)fa.chunk_rows;
    const int64_t left = M.m - r0;
    const int nrows = (int)(left < fa.chunk_rows ? left : fa.chunk_rows);
    double* gout_row = fa.gpartf + c * fa.npadf;
    const double* bvec = P.fvec;

    // ---- registers of the two role groups (as in fused_pass, variant 1) ------------------------------------------------
    double2 reg[kFH];                                 // dot warps: x; update warps: gradient accumulators (one array: 64 registers)
    double2 (&xr)[kFH] = reg;
    double2 (&acc)[kFH] = reg;
    const int tg = (warp < kFGWarps) ? t : t - kFGroup;
    if (warp < kFGWarps) {
      const int64_t n = M.n;
#pragma unroll
      for (int k = 0; k < kFH; ++k) {
        const int64_t j = col0 + 2 * (k * kFGroup + tg);
        xr[k].x = (j < n) ? ldcg(x + j) : 0.0;
        xr[k].y = (j + 1 < n) ? ldcg(x + j + 1) : 0.0;
      }
    } else {
#pragma unroll
      for (int k = 0; k < kFH; ++k) acc[k] = make_double2(0.0, 0.0);
    }
    double fsum = 0.0;
    const uint32_t tile0 = ring + tg * 16;

    for (int rb = 0; rb < nrows; rb += R) {
      const int nb = (nrows - rb < R) ? nrows - rb : R;
      double* xcur = xch + (size_t)(batch_no & 1u) * (size_t)R * (kFMaxCluster * kFGWarps);
      const char* src0 = reinterpret_cast<const char*>(M.a + col0 + (r0 + rb) * M.ld);
      // ======== pass A: partial dots of the batch rows ========
      if (warp < kFGWarps) {
        uint32_t gg = g;
        for (int i = 0; i < nb; ++i) {
          const uint32_t slot = gg % kFStages, ph = (gg / kFStages) & 1u;
          group_wait(warp == 0, 1, full0 + 8 * slot, ph);
          const uint32_t tile = tile0 + slot * kFStageBytes;
          double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
          double2 av[kFB];
#pragma unroll
          for (int h = 0; h < kFH; h += kFB) {
#pragma unroll
            for (int k = 0; k < kFB; ++k) av[k] = lds2v(tile + (h + k) * kFGroup * 16);
#pragma unroll
            for (int k = kFB - 2; k >= 0; k -= 2) {
              p2 = fma(av[k + 1].x, xr[h + k + 1].x, p2);
              p3 = fma(av[k + 1].y, xr[h + k + 1].y, p3);
              p0 = fma(av[k].x, xr[h + k].x, p0);
              p1 = fma(av[k].y, xr[h + k].y, p1);
            }
          }
          const double pw = warp_sum((p0 + p1) + (p2 + p3));                   // the cluster kernel's tree, bit for bit
          if (lane == 0) {
            xcur[(size_t)i * (kFMaxCluster * kFGWarps) + rho * kFGWarps + warp] = pw;
            mbar_arrive(empty0 + 8 * slot);
          }
          gg = (gg + 1) % 6u;
        }
      } else if (tg == 0) {
        // producer of pass A: lane 0 of the first update warp (the update warps have nothing else to do here)
        uint32_t gg = g;
        const char* src = src0;
        const uint64_t pol = l2_policy_evict_last();      // these rows come back in pass B
        for (int i = 0; i < nb; ++i) {
          const uint32_t slot = gg % kFStages, ph = (gg / kFStages) & 1u;
          fmbar_wait_hint(empty0 + 8 * slot, ph ^ 1u);
          mbar_expect_tx(full0 + 8 * slot, bytes);
          bulk_g2s_hint(ring + slot * kFStageBytes, src, bytes, full0 + 8 * slot, pol);
          src += ldb;
          gg = (gg + 1) % 6u;
        }
      }
      g = (g + (uint32_t)nb) % 6u;
      vb.sync();                                      // every CTA's partials of the batch are in global memory
      // ======== r_i of the batch: the 16 x 8 partials in the cluster kernel's order ========
      for (int i = warp; i < nb; i += kFWarps) {
        const double* pp = xcur + (size_t)i * (kFMaxCluster * kFGWarps);
        const double v0 = (lane < nval) ? ldcg(pp + lane) : 0.0;
        const double v1 = (lane + 32 < nval) ? ldcg(pp + lane + 32) : 0.0;
        const double v2 = (lane + 64 < nval) ? ldcg(pp + lane + 64) : 0.0;
        const double v3 = (lane + 96 < nval) ? ldcg(pp + lane + 96) : 0.0;
        const double rs = warp_sum((v0 + v1) + (v2 + v3)) - __ldg(bvec + r0 + rb + i);      // lasso/runme.jl:22
        if (lane == 0) s_r[i] = rs;
      }
      __syncthreads();
      // ======== pass B: rank-1 updates from the re-streamed rows ========
      if (warp >= kFGWarps) {
        uint32_t gg = g;
        for (int i = 0; i < nb; ++i) {
          const uint32_t slot = gg % kFStages, ph = (gg / kFStages) & 1u;
          group_wait(warp == kFGWarps, 2, full0 + 8 * slot, ph);
          const double rs = s_r[i];
          const uint32_t tile = tile0 + slot * kFStageBytes;
          double2 av[kFB];
#pragma unroll
          for (int h = 0; h < kFH; h += kFB) {
#pragma unroll
            for (int k = 0; k < kFB; ++k) av[k] = lds2v(tile + (h + k) * kFGroup * 16);
            __syncwarp();
#pragma unroll
            for (int k = kFB - 1; k >= 0; --k) {
              acc[h + k].x = fma(av[k].x, rs, acc[h + k].x);
              acc[h + k].y = fma(av[k].y, rs, acc[h + k].y);
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(empty0 + 8 * slot);
          fsum = fma(rs, rs, fsum);
          gg = (gg + 1) % 6u;
        }
      } else if (t == 0) {
        // producer of pass B: lane 0 of the first dot warp
        uint32_t gg = g;
        const char* src = src0;
        const uint64_t pol = l2_policy_evict_first();     // last use
        for (int i = 0; i < nb; ++i) {
          const uint32_t slot = gg % kFStages, ph = (gg / kFStages) & 1u;
          fmbar_wait_hint(empty0 + 8 * slot, ph ^ 1u);
          mbar_expect_tx(full0 + 8 * slot, bytes);
          bulk_g2s_hint(ring + slot * kFStageBytes, src, bytes, full0 + 8 * slot, pol);
          src += ldb;
          gg = (gg + 1) % 6u;
        }
      }
      g = (g + (uint32_t)nb) % 6u;
      ++batch_no;
      __syncthreads();                                // s_r may be overwritten by the next batch
    }
    // ---- chunk outputs: the same bits the cluster kernel leaves ---------------------------------------------------------
    if (warp >= kFGWarps) {
      double* gout = gout_row + col0;
#pragma unroll
      for (int k = 0; k < kFH; ++k) *reinterpret_cast<double2*>(gout + 2 * (k * kFGroup + tg)) = acc[k];
      if (tg == 0 && rho == 0) fa.fpart[c] = fsum;
    }
    vb.sync();                                        // (fence + arrival of every CTA after its stores)
    if (rho == 0 && t == 0) { __threadfence(); atomicAdd(fa.done, 1ull); atomicAdd(fa.hsync + 10, 1ull); }   // [10]: chunks the helpers processed (statistics)
  }
}

#endif  // ADAPROX_FUSED_VARIANT == 1

}  // namespace adaprox
