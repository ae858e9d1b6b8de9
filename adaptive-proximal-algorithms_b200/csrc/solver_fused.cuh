// solver_fused.cuh -- AdaPGM on a dense least-squares term with ONE sweep over A per
// iteration:  g = A'(A x - b)  (lasso/runme.jl:21-25 value + pullback, fused).
//
// The two-pass formulation streams A twice (A*x needs whole rows before A'r can start).
// Here a thread-block CLUSTER of C = ceil(ld / 8192) <= 16 CTAs owns whole rows
// (8192 columns per CTA).  For every row i:
//   1. one lane per CTA (the producer) has bulk-copied the CTA's 64 KB piece of row i into a
//      3-slot shared-memory ring (cp.async.bulk + mbarrier);
//   2. the 8 DOT warps (x in registers) form the partial dot <A[i, cols], x[cols]> from shared
//      memory and push one partial per warp into every peer's shared memory with
//      st.async ... mbarrier::complete_tx (DSMEM), 8-deep exchange buffers;
//   3. the 8 UPDATE warps (column accumulators in registers) wait for the C * 8 partials of the
//      row, sum them in a fixed order (same bits in every warp of every CTA of the cluster),
//      r[i] = sum - b[i], f += r[i]^2, and apply acc[cols] += A[i, cols] * r[i] from the tile
//      that is still resident; the freed slot is refilled with row i + 3.
//   Waits: one warp per role group polls the mbarrier, the other seven park on a named
//   barrier (sleeping waiters -- try_wait, nanosleep -- halve the speed of the sweep).
// Rows are handed to the clusters in chunks, dynamically; every chunk stores its own partial
// (gpartf[chunk][n], fpart[chunk]) and the chunks are reduced in chunk order, so results do not
// depend on the schedule and reruns are bit-identical.  DRAM traffic per iteration is 8 m n
// instead of 16 m n.  Kernels: k_adapgm_fused = the whole single-GPU solve as one cluster
// launch with an own grid barrier; its sweep-only mode serves the row-sharded solve
// (comm.inl), optionally with the all-reduce inside the kernel (p2p.cuh).
#pragma once
#include "phases.cuh"
#include "p2p.cuh"

namespace adaprox {

constexpr int kFGroup = 256;                              // threads per role group (dot warps / update warps)
constexpr int kFGWarps = kFGroup / 32;
constexpr int kFThreads = 2 * kFGroup;                    // 8 dot warps + 8 update warps (512 threads: 128 registers each)
constexpr int kFWarps = kFThreads / 32;
constexpr int kFCols = 8192;                              // columns per CTA
constexpr int kFH = kFCols / 2 / kFGroup;                 // 16 double2 per thread per row
constexpr int kFStages = 3;
constexpr int kFStageBytes = kFCols * 8;                  // 64 KB
constexpr int kFRingBytes = kFStages * kFStageBytes;      // 192 KB dynamic shared memory
constexpr int kFMaxCluster = 16;
// Two implementations of the row pipeline (build flag -DADAPROX_FUSED_VARIANT=1|2; measured in profiles/r02_notes.md):
//   1 (shipped): role-specialised warps -- 8 dot warps and 8 update warps both read the row tile from shared memory; the
//      two latency chains overlap.  Throughput is bounded by Little's law on the 3 x 64 KB ring: a slot is busy for
//      copy (2700 cycles under load) + dot (1100) + exchange (900) + update (1500), period = lifetime / 3 = 2150 cycles.
//   2 (A/B only): every warp does both jobs on its own 16 columns with ONE shared-memory pass per row into registers and
//      releases the slot right after it.  Same results, but x + accumulators + one row fill the register file, so dot,
//      exchange and update of a row cannot overlap with the next row: 3000 cycles per row (14.1 vs 11.4 ms per sweep).
#ifndef ADAPROX_FUSED_VARIANT
#define ADAPROX_FUSED_VARIANT 1
#endif
#if ADAPROX_FUSED_VARIANT == 1
constexpr int kFDepth = 8;                                // exchange-buffer depth (see fused_pass)
constexpr int kFXWarps = kFGWarps;                        // warps per CTA that contribute a partial dot per row
constexpr int kFsumThread = kFGroup;                      // the thread of cluster rank 0 that returns sum r_i^2
#else
constexpr int kFDepth = 4;
constexpr int kFXWarps = kFWarps;
constexpr int kFsumThread = 0;
constexpr int kFH2 = kFCols / 2 / kFThreads;              // 8 double2 per thread per row
#endif
constexpr int kFTraceRows = 512;
#ifdef ADAPROX_FUSED_TRACE
#define FTRACE(cond, k, i) do { if ((cond) && tr != nullptr && (i) < kFTraceRows) tr[(k) * kFTraceRows + (i)] = (unsigned long long)clock64(); } while (0)
#else
#define FTRACE(cond, k, i) do { } while (0)
#endif

struct FusedArgs {
  double* gpartf;      // [nchunks][npadf] per-chunk A'r partials
  double* fpart;       // [nchunks] per-chunk sums of r_i^2
  int64_t npadf;       // C * 8192
  int C;               // cluster size
  int chunk_rows;      // rows per chunk (unit of the dynamic schedule)
  int nchunks;         // ceil(m / chunk_rows)
  unsigned long long* bar;       // grid-barrier arrival counter (zeroed by the host before the launch)
  unsigned long long bar_base;   // arrivals already counted on `bar` by earlier launches (row-sharded solve)
  unsigned long long* next;      // chunk dispenser (monotonic; zeroed by the host before the launch)
  unsigned long long next_base;  // dispenser value at the start of this launch's first sweep
  // sweep-only mode (row-sharded solve, comm.inl): one sweep for sh_x, shard partials into sh_gbuf[n + 2], exit
  int sweep_only;
  const double* sh_x;
  double* sh_gbuf;
  const int* sh_done;            // the solve has stopped: do nothing
  P2PArgs p2p;                   // in-kernel all-reduce over peer-mapped buffers (p2p.cuh); p2p.n <= 1: ncclAllReduce by the host
  // Helper CTAs on the SMs no 16-CTA cluster can occupy (k_adapgm_helper, persistent mode only; hV = 0: none).  They take row
  // chunks from the same dispenser and leave the same per-chunk partials, so the result does not depend on who processed what.
  int tagged;                    // dispenser protocol: value = (sweep << 32) | index, reset by CTA 0 before the barrier that precedes a sweep
  int hV, hR;                    // virtual clusters of 16 helper CTAs; rows per helper batch
  int hHold;                     // helpers leave the last hHold chunks of a sweep to the clusters (a helper chunk takes ~1.7x as long: no tail)
  unsigned long long* go;        // sweep whose iterate is complete (written after that barrier); helpers wait for go >= tag
  unsigned long long* done;      // chunks completed so far, all sweeps, by main clusters and helpers alike
  unsigned long long* hsync;     // [0] roll call, [1] abort, [2 .. 2 + hV) barrier counters, [2 + hV .. 2 + 2 hV) chunk mailboxes of the virtual clusters
  double* hxch;                  // [hV][2][hR][16 * 8] partial dots of a helper batch (global memory instead of DSMEM)
  unsigned long long* resident;  // mapped host word: the cluster kernel writes resident_seq once all its CTAs run (gate of the helper launch)
  unsigned long long resident_seq;
  int* err;                      // set by a CTA whose grid barrier timed out (GridBar); checked by the host after the solve
  unsigned long long* lat;       // optional [grid][4] probe (ADAPROX_FUSED_LAT): chunks taken, sweep ns, -, smid
  unsigned long long* trace;     // build flag ADAPROX_FUSED_TRACE only: [5][kFTraceRows] clock64 stamps of CTA 0 (issue, full, dot done, exchange complete, update done)
};

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t ncluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// Control block in static shared memory: full[3] | empty[3] | cfull[8] mbarriers, then the exchange buffers
// cpart[8][16 ranks][8 warps].  One base register + compile-time offsets address all of it.
constexpr uint32_t kOffFull = 0, kOffEmpty = 8 * kFStages, kOffCfull = 16 * kFStages, kOffChunk = 16 * kFStages + 8 * kFDepth, kOffCpart = 128;
constexpr uint32_t kPartStride = kFMaxCluster * kFXWarps * 8;                 // bytes per exchange buffer
constexpr uint32_t kOffRs = kOffCpart + kFDepth * kPartStride;                // r_i of the row in exchange buffer d (variant 1, one reducer warp)
constexpr int kCtlBytes = kOffRs + kFDepth * 8;
static_assert(kOffChunk + 8 <= kOffCpart, "control block layout");

struct FusedSmem {
  uint32_t ring, ctl;                         // shared-space addresses of the tile ring and the control block
  uint32_t count;                             // rows pushed through the ring / exchange so far (uniform)
};

// remote 8-byte store that completes 8 bytes on the destination CTA's mbarrier (no fence needed on the sender)
__device__ __forceinline__ void st_async_peer(uint32_t local_data_addr, uint32_t local_mbar_addr, uint32_t peer, double v) {
  uint32_t rdata, rbar;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rdata) : "r"(local_data_addr), "r"(peer));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(local_mbar_addr), "r"(peer));
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
               ::"r"(rdata), "l"(__double_as_longlong(v)), "r"(rbar) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  } while (!ok);
}

// Waits of the fused pass POLL (mbarrier.test_wait in a loop) instead of using the suspending try_wait.  Measured on
// the same B200, same build otherwise, 65536 x 131072: 11.99 ms per sweep with polling, 27.98 ms with try_wait.  A
// warp that try_wait suspended is resumed late; in this pipeline every late consumer delays the refill of its ring
// slot, which makes the next wait block as well -- once a cluster drops into that regime it stays there (all warps
// on the long scoreboard, DRAM at 37 %).  With only 16 warps per SM the issue slots the polling costs are free.
// ADAPROX_FUSED_TRYWAIT (build flag) restores the suspending wait for A/B measurements.
#ifndef ADAPROX_FUSED_TRYWAIT
__device__ __forceinline__ void fmbar_wait(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  for (;;) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (ok) break;
#ifdef ADAPROX_FUSED_POLL_SLEEP_NS
    __nanosleep(ADAPROX_FUSED_POLL_SLEEP_NS);
#endif
  }
}
#else
__device__ __forceinline__ void fmbar_wait(uint32_t addr, uint32_t parity) { mbar_wait(addr, parity); }
#endif

// One warp of a role group polls, the other seven park on a hardware named barrier (bar.sync id, 256): parked warps issue
// nothing (unlike 16 polling warps, which keep the issue slots and the power budget busy) and are released within a few
// cycles of the poller's arrival (unlike warps suspended by try_wait).  Build flag ADAPROX_FUSED_ALLPOLL restores "every warp polls".
#if ADAPROX_FUSED_VARIANT == 1
constexpr int kFBarThreads = kFGroup;                     // a role group
#else
constexpr int kFBarThreads = kFThreads;                   // the whole CTA
#endif
__device__ __forceinline__ void group_bar(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kFBarThreads) : "memory"); }
#ifndef ADAPROX_FUSED_ALLPOLL
#ifndef ADAPROX_FUSED_TRYWAIT_NS
#define ADAPROX_FUSED_TRYWAIT_NS 64
#endif
#if ADAPROX_FUSED_TRYWAIT_NS <= 0
__device__ __forceinline__ void fmbar_wait_hint(uint32_t addr, uint32_t parity) { fmbar_wait(addr, parity); }
#else
// The poller of a role group suspends in hardware (mbarrier.try_wait with a SHORT time hint) instead of spinning on test_wait.
// Alone it changes nothing (84.9 vs 84.5 it/s; hints of 32 ... 5000 ns alike) -- unlike round 1's try_wait WITHOUT a hint,
// which was 2.3x slower -- but the sweep kernel runs into the 1000 W power cap, and with suspended pollers the helper CTAs turn
// the freed power into throughput (93.7 vs 84.5 it/s); with spinning pollers they gained nothing.  -DADAPROX_FUSED_TRYWAIT_NS=0
// restores the spinning poller.
__device__ __forceinline__ void fmbar_wait_hint(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(addr), "r"(parity), "r"((uint32_t)ADAPROX_FUSED_TRYWAIT_NS) : "memory");
  } while (!ok);
}
#endif
__device__ __forceinline__ void group_wait(bool poller, int id, uint32_t addr, uint32_t parity) {
#if ADAPROX_FUSED_TRYWAIT_NS > 0
  if (poller) fmbar_wait_hint(addr, parity);
#elif defined(ADAPROX_FUSED_POLL_LANE0)
  // experiment: one LANE polls (the other 31 lanes of the polling warp wait at the warp barrier): fewer active lanes per poll
  if (poller) { if ((threadIdx.x & 31) == 0) fmbar_wait(addr, parity); __syncwarp(); }
#else
  if (poller) fmbar_wait(addr, parity);
#endif
  group_bar(id);
}
#else
__device__ __forceinline__ void group_wait(bool, int, uint32_t addr, uint32_t parity) { fmbar_wait(addr, parity); }
#endif

// ld.volatile keeps the program order of the loads; the consumers below use the batch in REVERSE order, so all
// eight loads of a batch must be in flight before the first FMA can issue (ptxas otherwise recycles one
// destination quad: load -> FMA -> load ..., one shared-memory round trip per 16 bytes).
__device__ __forceinline__ double2 lds2v(uint32_t addr) {
  double2 v;
  asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}

// fused_pass: rows [r0, r0 + nrows) of the matrix.  gout_row[cols] = sum_i A[i, cols] * (A[i,:] x - b[i]);
// returns sum r_i^2 (non-zero in one thread of cluster rank 0 only).
//
// Warp roles (the groups are coupled through mbarriers only, so the latency chains of one overlap the
// shared-memory bursts of the other):
//   producer (lane 0 of the last update warp): bulk copies row g + 3 into the ring slot row g just left;
//   dot warps (0..7): x in registers; per row 16 x LDS.128 + 32 DFMA per thread, warp sum, then lanes < C push
//     the warp's partial into cpart[g % 8][rank][warp] of every peer with st.async, which completes 8 bytes on
//     the peer's cfull[g % 8] mbarrier;
//   update warps (8..15): accumulators in registers; wait for the C * 8 partials of row g, sum them in a fixed
//     order (bit-identical in every warp of every CTA of the cluster), r = sum - b, rank-1 update from the tile
//     still resident in shared memory, free the slot.
// Exchange-buffer depth: a CTA can push row j only after its own update of row j-3 (ring slot), which needs
// every peer's dot of row j-3, which needs that peer's update of row j-6: when row j's partials arrive, a peer
// may still be reading rows j-5 .. j-1, so 8 buffers never collide (and phase j-8 of cfull is long complete).
// Register budget: 512 threads leave 128 registers per thread, so the loop state is kept minimal: x or the
// accumulators (64 registers), one batch of tile data (16), a handful of 32-bit addresses and counters.
#ifndef ADAPROX_FUSED_KFB
#define ADAPROX_FUSED_KFB 4
#endif
constexpr int kFB = ADAPROX_FUSED_KFB;                     // LDS.128 per batch (16 registers of tile data per thread)

#if ADAPROX_FUSED_VARIANT == 1
__device__ __noinline__ double fused_pass(const DMat& M, const double* bvec, const double* x, FusedSmem& fs, int C,
                                           int64_t r0, int nrows, double* gout_row, unsigned long long* tr = nullptr) {
  const uint32_t ring = fs.ring, ctl = fs.ctl, g0 = fs.count;
  (void)tr;
  const uint32_t rank = cluster_ctarank();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t col0 = (int64_t)rank * kFCols;
  double fsum = 0.0;

  if (warp < kFGWarps) {
    // ------------------------------------------------------------------ dot warps
    const int t = threadIdx.x;
    double2 xr[kFH];
    {
      const int64_t n = M.n;
#pragma unroll
      for (int k = 0; k < kFH; ++k) {
        const int64_t j = col0 + 2 * (k * kFGroup + t);
        xr[k].x = (j < n) ? ldcg(x + j) : 0.0;
        xr[k].y = (j + 1 < n) ? ldcg(x + j + 1) : 0.0;
      }
    }
    uint32_t slot = g0 % kFStages, ph = (g0 / kFStages) & 1u, d = g0 % kFDepth;
    const uint32_t tile0 = ring + t * 16;
    const uint32_t mypart = ctl + kOffCpart + (rank * kFGWarps + warp) * 8;
    const bool sender = lane < C;
    // remote addresses of this warp's slot in peer `lane`: mapped once, buffer d is a constant offset (the kernel is power-limited:
    // every instruction taken out of the per-row loop counts, profiles/r02_notes.md section 7)
    uint32_t rdata0 = 0, rbar0 = 0;
    if (sender) {
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rdata0) : "r"(mypart), "r"((uint32_t)lane));
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar0) : "r"(ctl + kOffCfull), "r"((uint32_t)lane));
    }
    for (int i = 0; i < nrows; ++i) {
      group_wait(warp == 0, 1, ctl + kOffFull + 8 * slot, ph);
      FTRACE(t == 0, 1, i);
      const uint32_t tile = tile0 + slot * kFStageBytes;
      // batches of kFB x LDS.128 issued back to back: ld.volatile keeps their order and the FMA chains consume the
      // batch in REVERSE, so the whole batch is in flight before the first FMA (ptxas otherwise recycles ONE
      // destination quad: load -> FMA -> load ..., a full shared-memory round trip per 16 bytes)
      double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
      double2 av[kFB];
#pragma unroll
      for (int h = 0; h < kFH; h += kFB) {
#pragma unroll
#ifdef ADAPROX_EXP_NO_DOT_LDS
        for (int k = 0; k < kFB; ++k) av[k] = xr[(h + k + 1) % kFH];          // timing experiment only: no shared-memory reads in the dot warps
#else
        for (int k = 0; k < kFB; ++k) av[k] = lds2v(tile + (h + k) * kFGroup * 16);
#endif
#pragma unroll
        for (int k = kFB - 2; k >= 0; k -= 2) {
          p2 = fma(av[k + 1].x, xr[h + k + 1].x, p2);
          p3 = fma(av[k + 1].y, xr[h + k + 1].y, p3);
          p0 = fma(av[k].x, xr[h + k].x, p0);
          p1 = fma(av[k].y, xr[h + k].y, p1);
        }
      }
      const double pw = warp_sum((p0 + p1) + (p2 + p3));
      if (sender)
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
                     ::"r"(rdata0 + d * kPartStride), "l"(__double_as_longlong(pw)), "r"(rbar0 + 8 * d) : "memory");
      FTRACE(t == 0, 2, i);
      if (++slot == kFStages) { slot = 0; ph ^= 1u; }
      d = (d + 1) % kFDepth;
    }
  } else {
    // ------------------------------------------------------------------ update warps
    const int t = threadIdx.x - kFGroup;
    const bool leader = (t == 0);
    const bool producer = (t == kFGroup - 32);       // lane 0 of the last update warp keeps the ring full
    const int nval = C * kFGWarps;                   // partials per row (<= 128)
    const uint32_t xbytes = (uint32_t)nval * 8;      // exchange bytes per row per CTA
    uint32_t bytes;
    {
      int64_t width = M.ld - col0;
      width = width < 0 ? 0 : (width > kFCols ? kFCols : width);
      bytes = (uint32_t)(width * 8);
    }
    const int64_t ldb = M.ld * 8;
    const char* src = reinterpret_cast<const char*>(M.a + col0 + r0 * M.ld);     // next row to copy (producer)
    double2 acc[kFH];
#pragma unroll
    for (int k = 0; k < kFH; ++k) acc[k] = make_double2(0.0, 0.0);
    if (leader)
      for (int i = 0; i < kFDepth && i < nrows; ++i) mbar_expect_tx(ctl + kOffCfull + 8 * ((g0 + i) % kFDepth), xbytes);
    uint32_t slot = g0 % kFStages, ph = (g0 / kFStages) & 1u, d = g0 % kFDepth, dph = (g0 / kFDepth) & 1u;
    int issued = 0; (void)issued;
    const uint64_t pol = l2_policy_evict_first();     // the matrix is read once per sweep: do not let it displace what the helper CTAs re-read
    auto issue = [&](uint32_t sl, uint32_t par) {    // wait until the update warps left slot sl, then refill it
      fmbar_wait(ctl + kOffEmpty + 8 * sl, par ^ 1u);
      FTRACE(true, 0, issued); ++issued;
      mbar_expect_tx(ctl + kOffFull + 8 * sl, bytes);
      bulk_g2s_hint(ring + sl * kFStageBytes, src, bytes, ctl + kOffFull + 8 * sl, pol);
      src += ldb;
    };
    if (producer) {
      uint32_t sl = slot, par = ph;
      for (int i = 0; i < kFStages && i < nrows; ++i) {
        issue(sl, par);
        if (++sl == kFStages) { sl = 0; par ^= 1u; }
      }
    }
    const uint32_t tile0 = ring + t * 16;
    const uint32_t part0 = ctl + kOffCpart + lane * 8;
    const double* bp = bvec + r0;
    double bblk = 0.0;                               // lane l holds b[r0 + 32 * (i / 32) + l]
    for (int i = 0; i < nrows; ++i) {
#if !defined(ADAPROX_FUSED_ALLPOLL) && !defined(ADAPROX_FUSED_ALL_REDUCE)
      // ONE warp per CTA (the poller of the update group) waits for the exchange, sums the C * 8 partials in the fixed order (the same bits
      // in every CTA of the cluster) and publishes r_i through shared memory before it joins the group barrier: the other seven warps
      // were executing the identical reduction (22 instructions each per row, ~11 % of the kernel's instruction stream -- and the kernel
      // is power-limited, profiles/r02_notes.md section 7).  -DADAPROX_FUSED_ALL_REDUCE restores the redundant form.
      if (warp == kFGWarps) {
        if ((i & 31) == 0) bblk = (i + lane < nrows) ? __ldg(bp + i + lane) : 0.0;
        const double b_cur = __shfl_sync(0xffffffffu, bblk, i & 31);
        // st.async delivers data and complete_tx through the same path into this CTA's shared memory, so the
        // cta-scope wait is enough (a cluster-scope acquire compiles to CCTL.IVALL: an L1 flush per row)
        fmbar_wait_hint(ctl + kOffCfull + 8 * d, dph);
        const uint32_t pb = part0 + d * kPartStride;
        const double v0 = (lane < nval) ? lds1(pb) : 0.0;
        const double v1 = (lane + 32 < nval) ? lds1(pb + 256) : 0.0;
        const double v2 = (lane + 64 < nval) ? lds1(pb + 512) : 0.0;
        const double v3 = (lane + 96 < nval) ? lds1(pb + 768) : 0.0;
        if (leader && i + kFDepth < nrows) mbar_expect_tx(ctl + kOffCfull + 8 * d, xbytes);    // arm this buffer for row g + 8
        const double r1 = warp_sum((v0 + v1) + (v2 + v3)) - b_cur;                            // lasso/runme.jl:22  res = A*w - b
        if (lane == 0) sts1(ctl + kOffRs + 8 * d, r1);
      }
      group_bar(2);                                  // (bar.sync orders the reducer's shared-memory store before the reads below)
      FTRACE(leader, 3, i);
      const double rs = lds1(ctl + kOffRs + 8 * d);  // slot d is rewritten 8 rows -- 8 group barriers -- later
#else
      if ((i & 31) == 0) bblk = (i + lane < nrows) ? __ldg(bp + i + lane) : 0.0;
      const double b_cur = __shfl_sync(0xffffffffu, bblk, i & 31);
      // st.async delivers data and complete_tx through the same path into this CTA's shared memory, so the
      // cta-scope acquire is enough (a cluster-scope acquire compiles to CCTL.IVALL: an L1 flush per row)
      group_wait(warp == kFGWarps, 2, ctl + kOffCfull + 8 * d, dph);
      FTRACE(leader, 3, i);
      const uint32_t pb = part0 + d * kPartStride;
      const double v0 = (lane < nval) ? lds1(pb) : 0.0;
      const double v1 = (lane + 32 < nval) ? lds1(pb + 256) : 0.0;
      const double v2 = (lane + 64 < nval) ? lds1(pb + 512) : 0.0;
      const double v3 = (lane + 96 < nval) ? lds1(pb + 768) : 0.0;
      if (leader && i + kFDepth < nrows) mbar_expect_tx(ctl + kOffCfull + 8 * d, xbytes);    // arm this buffer for row g + 8
      const double rs = warp_sum((v0 + v1) + (v2 + v3)) - b_cur;   // same order in every update warp of the cluster;
                                                                   // lasso/runme.jl:22  res = A*w - b
#endif
#ifdef ADAPROX_FUSED_ALLPOLL
      fmbar_wait(ctl + kOffFull + 8 * slot, ph);                 // long complete; makes the bulk-copied tile visible here
#endif
      // (default build: the tile was observed complete by the dot warps' poller, whose partial sum is part of the
      //  exchange this warp was just released on -- mbarrier completion -> named barrier is a causality chain)
      const uint32_t tile = tile0 + slot * kFStageBytes;
      double2 av[kFB];
#pragma unroll
      for (int h = 0; h < kFH; h += kFB) {
#pragma unroll
#ifdef ADAPROX_EXP_NO_UPD_LDS
        for (int k = 0; k < kFB; ++k) av[k] = make_double2(b_cur, rs);          // timing experiment only: no shared-memory reads in the update warps
#else
        for (int k = 0; k < kFB; ++k) av[k] = lds2v(tile + (h + k) * kFGroup * 16);
#endif
#ifndef ADAPROX_EXP_NO_SYNCWARP
        __syncwarp();                                           // scheduling fence: the independent FMAs below must not be
                                                                // hoisted between the loads (one load in flight otherwise)
#endif
#pragma unroll
        for (int k = kFB - 1; k >= 0; --k) {
          acc[h + k].x = fma(av[k].x, rs, acc[h + k].x);
          acc[h + k].y = fma(av[k].y, rs, acc[h + k].y);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(ctl + kOffEmpty + 8 * slot);
      FTRACE(leader, 4, i);
      if (producer && i + kFStages < nrows) issue(slot, ph ^ 1u);   // row i + 3 into the slot row i just left
      if (warp == kFGWarps) fsum = fma(rs, rs, fsum);               // (only the leader's copy is returned)
      if (++slot == kFStages) { slot = 0; ph ^= 1u; }
      if (++d == kFDepth) { d = 0; dph ^= 1u; }
    }
    double* gout = gout_row + col0;
#pragma unroll
    for (int k = 0; k < kFH; ++k) *reinterpret_cast<double2*>(gout + 2 * (k * kFGroup + t)) = acc[k];
    if (!(leader && rank == 0)) fsum = 0.0;
  }
  __syncthreads();
  // slot = g % 3, ring phase = (g / 3) & 1, exchange buffer = g % 8, its phase = (g / 8) & 1: everything is periodic in g with
  // period lcm(6, 16) = 48, so the running row index is kept modulo 48 (a 32-bit count would wrap after 4e9 rows)
  fs.count = (uint32_t)(((uint64_t)g0 + (uint64_t)nrows) % 48u);
  return fsum;
}
#else
// Variant 2.  All 16 warps run the same code; thread t owns the double2 columns k * 512 + t (k = 0 .. 7) of the CTA's
// 8192-column slice: x (32 registers), the gradient accumulators (32) and ONE row of tile data (32) live in registers.
// Per row g:
//   1. warp 0 polls full[g % 3], the other warps park on a named barrier;
//   2. 8 x LDS.128 bring the thread's 16 columns into registers -- the only shared-memory read of the tile; once the
//      partial dot (which depends on all of them) exists, lane 0 of every warp arrives on empty[g % 3] and thread 0
//      refills the slot with row g + 3: a slot is busy for copy + one LDS pass instead of copy + dot + exchange + update;
//   3. warp partial -> st.async into cpart[g % 4][rank][warp] of every peer (complete_tx on the peer's cfull[g % 4]);
//   4. warp 1 polls cfull[g % 4]; all warps sum the C * 16 partials in a fixed order (the same bits in every warp of
//      every CTA of the cluster): r = sum - b[g];
//   5. acc += tile registers * r.
// Exchange depth: a warp sends row g + 1 only after it has read the partials of row g, and a peer needs ALL 16 warp
// partials of row g + 1 from this CTA before it can finish row g + 1 and send row g + 2 -- so buffer g % 4 is never
// written (row g + 4, or even g + 2) while a warp here still reads row g.  Everything is periodic in g with period
// lcm(2 * 3, 2 * 4) = 24.
__device__ __noinline__ double fused_pass(const DMat& M, const double* bvec, const double* x, FusedSmem& fs, int C,
                                           int64_t r0, int nrows, double* gout_row, unsigned long long* tr = nullptr) {
  const uint32_t ring = fs.ring, ctl = fs.ctl, g0 = fs.count;
  (void)tr;
  const uint32_t rank = cluster_ctarank();
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int64_t col0 = (int64_t)rank * kFCols;
  double fsum = 0.0;

  double2 xr[kFH2], acc[kFH2];
  {
    const int64_t n = M.n;
#pragma unroll
    for (int k = 0; k < kFH2; ++k) {
      const int64_t j = col0 + 2 * (k * kFThreads + t);
      xr[k].x = (j < n) ? ldcg(x + j) : 0.0;
      xr[k].y = (j + 1 < n) ? ldcg(x + j + 1) : 0.0;
      acc[k] = make_double2(0.0, 0.0);
    }
  }
  const int nval = C * kFXWarps;                     // partials per row (<= 256)
  const uint32_t xbytes = (uint32_t)nval * 8;        // exchange bytes per row per CTA
  uint32_t bytes;
  {
    int64_t width = M.ld - col0;
    width = width < 0 ? 0 : (width > kFCols ? kFCols : width);
    bytes = (uint32_t)(width * 8);
  }
  const int64_t ldb = M.ld * 8;
  const char* src = reinterpret_cast<const char*>(M.a + col0 + r0 * M.ld);     // next row to copy (thread 0)
  uint32_t slot = g0 % kFStages, ph = (g0 / kFStages) & 1u, d = g0 % kFDepth, dph = (g0 / kFDepth) & 1u;
  auto issue = [&](uint32_t sl, uint32_t par) {      // wait until every warp has left slot sl, then refill it
    fmbar_wait(ctl + kOffEmpty + 8 * sl, par ^ 1u);
    mbar_expect_tx(ctl + kOffFull + 8 * sl, bytes);
    bulk_g2s(ring + sl * kFStageBytes, src, bytes, ctl + kOffFull + 8 * sl);
    src += ldb;
  };
  if (t == 0) {
    for (int i = 0; i < kFDepth && i < nrows; ++i) mbar_expect_tx(ctl + kOffCfull + 8 * ((g0 + i) % kFDepth), xbytes);
    uint32_t sl = slot, par = ph;
    for (int i = 0; i < kFStages && i < nrows; ++i) {
      FTRACE(true, 0, i);
      issue(sl, par);
      if (++sl == kFStages) { sl = 0; par ^= 1u; }
    }
  }
  const uint32_t tile0 = ring + t * 16;
  const uint32_t mypart = ctl + kOffCpart + (rank * kFXWarps + warp) * 8;
  const uint32_t part0 = ctl + kOffCpart + lane * 8;
  const bool sender = lane < C;
  const double* bp = bvec + r0;
  double bblk = 0.0;                                 // lane l holds b[r0 + 32 * (i / 32) + l]
  for (int i = 0; i < nrows; ++i) {
    if ((i & 31) == 0) bblk = (i + lane < nrows) ? __ldg(bp + i + lane) : 0.0;
    const double b_cur = __shfl_sync(0xffffffffu, bblk, i & 31);
    group_wait(warp == 0, 1, ctl + kOffFull + 8 * slot, ph);
    FTRACE(t == 0, 1, i);
    const uint32_t tile = tile0 + slot * kFStageBytes;
    double2 av[kFH2];
#pragma unroll
    for (int k = 0; k < kFH2; ++k) av[k] = lds2v(tile + k * kFThreads * 16);
    // the chains consume the batch in REVERSE order: all eight loads are in flight before the first FMA can issue
    double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
#pragma unroll
    for (int k = kFH2 - 2; k >= 0; k -= 2) {
      p2 = fma(av[k + 1].x, xr[k + 1].x, p2);
      p3 = fma(av[k + 1].y, xr[k + 1].y, p3);
      p0 = fma(av[k].x, xr[k].x, p0);
      p1 = fma(av[k].y, xr[k].y, p1);
    }
    const double pw = warp_sum((p0 + p1) + (p2 + p3));     // depends on all eight loads: the slot is no longer needed
    if (lane == 0) mbar_arrive(ctl + kOffEmpty + 8 * slot);
    if (sender) st_async_peer(mypart + d * kPartStride, ctl + kOffCfull + 8 * d, (uint32_t)lane, pw);
    FTRACE(t == 0, 2, i);
    if (t == 0 && i + kFStages < nrows) { FTRACE(true, 0, i + kFStages); issue(slot, ph ^ 1u); }   // row i + 3 into the slot row i just left
    // st.async delivers data and complete_tx through the same path into this CTA's shared memory, so the
    // cta-scope wait is enough (a cluster-scope acquire compiles to CCTL.IVALL: an L1 flush per row)
    group_wait(warp == 1, 2, ctl + kOffCfull + 8 * d, dph);
    FTRACE(t == 0, 3, i);
    const uint32_t pb = part0 + d * kPartStride;
    double v[8];                                     // entries of ranks >= C are never written: zeroed once in fused_smem_init
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = lds1(pb + 256 * k);
    const double rs = warp_sum(((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]))) - b_cur;   // lasso/runme.jl:22  res = A*w - b
    __syncwarp();
    if (t == 0 && i + kFDepth < nrows) mbar_expect_tx(ctl + kOffCfull + 8 * d, xbytes);    // arm this buffer for row g + 4
#pragma unroll
    for (int k = 0; k < kFH2; ++k) {
      acc[k].x = fma(av[k].x, rs, acc[k].x);
      acc[k].y = fma(av[k].y, rs, acc[k].y);
    }
    FTRACE(t == 0, 4, i);
    fsum = fma(rs, rs, fsum);
    if (++slot == kFStages) { slot = 0; ph ^= 1u; }
    if (++d == kFDepth) { d = 0; dph ^= 1u; }
  }
  double* gout = gout_row + col0;
#pragma unroll
  for (int k = 0; k < kFH2; ++k) *reinterpret_cast<double2*>(gout + 2 * (k * kFThreads + t)) = acc[k];
  if (!(t == kFsumThread && rank == 0)) fsum = 0.0;
  __syncthreads();
  fs.count = (uint32_t)(((uint64_t)g0 + (uint64_t)nrows) % 24u);
  return fsum;
}
#endif

// One sweep over the matrix: g = A'(A x - b) as per-chunk partials gpartf[c][cols], fpart[c] = sum of r_i^2 over
// the rows of chunk c.  Chunks of fa.chunk_rows rows are handed to the clusters DYNAMICALLY (one atomicAdd per
// chunk by the cluster's rank-0 CTA, the index mailed to every peer through DSMEM): the 16 CTAs of a cluster run
// in lock step, so a single SM with a slower path to L2 slows its whole cluster down, and with a static row split
// the slowest of the 7 clusters (often 1.6-2x behind on this part) set the time of the sweep.  The OUTPUT does
// not depend on who processed what: partials are stored per chunk and reduced in chunk order afterwards, so
// results stay bit-reproducible.  `sweep` = number of sweeps this launch has done before (dispenser base).
__device__ __forceinline__ void fused_sweep(const DMat& M, const double* bvec, const double* x, FusedSmem& fs, const FusedArgs& fa,
                                            unsigned long long sweep) {
  const uint32_t rank = cluster_ctarank();
  const unsigned long long base = fa.next_base + sweep * (unsigned long long)(fa.nchunks + (int)ncluster_id_x());
  const uint32_t mailbox = fs.ctl + kOffChunk;
  int taken = 0;
  for (;;) {
    if (rank == 0 && threadIdx.x == 0) {
      long long c;
      if (fa.tagged) {
        c = (long long)(atomicAdd(fa.next, 1ull) & 0xffffffffull);           // tag = this sweep: reset before the preceding grid barrier
      } else {
        c = (long long)(atomicAdd(fa.next, 1ull) - base);
      }
      for (int p = 0; p < fa.C; ++p) {
        uint32_t raddr;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(mailbox), "r"(p));
        asm volatile("st.shared::cluster.s64 [%0], %1;" ::"r"(raddr), "l"(c) : "memory");
      }
    }
    cluster_arrive();          // release: the mailbox stores; acquire: every thread of every CTA sees its mailbox
    cluster_wait();
    // ... and every peer's gradient stores of the previous chunk are ordered before this point: publish its completion
    if (fa.tagged && taken > 0 && rank == 0 && threadIdx.x == 0) { __threadfence(); atomicAdd(fa.done, 1ull); }
    long long c;
    asm volatile("ld.volatile.shared.s64 %0, [%1];" : "=l"(c) : "r"(mailbox));
    if (c >= fa.nchunks) break;
    const int64_t r0 = c * (int64_t)fa.chunk_rows;
    const int64_t left = M.m - r0;
    const int nrows = (int)(left < fa.chunk_rows ? left : fa.chunk_rows);
#ifdef ADAPROX_FUSED_TRACE
    const double fv = fused_pass(M, bvec, x, fs, fa.C, r0, nrows, fa.gpartf + c * fa.npadf, (blockIdx.x == 0 && taken == 1) ? fa.trace : nullptr);
#else
    const double fv = fused_pass(M, bvec, x, fs, fa.C, r0, nrows, fa.gpartf + c * fa.npadf);
#endif
    if (rank == 0 && threadIdx.x == kFsumThread) fa.fpart[c] = fv;   // the one thread that holds the sum
    ++taken;
    // the next mailbox store happens after this CTA's rank-0 peer finished the chunk, which needs every CTA's
    // last partial dot, which every CTA sends after it has read the mailbox above: no overwrite race
  }
  if (fa.lat && threadIdx.x == 0) fa.lat[4 * blockIdx.x] += (unsigned long long)taken;
}

// grad[j] = sum over chunks in chunk order (fixed), for j in [j0, j1)
__device__ __forceinline__ void fused_gradient_slice(const FusedArgs& fa, int64_t j0, int64_t j1, double* out) {
  for (int64_t j = j0 + threadIdx.x; j < j1; j += kFThreads) {
    const double* p = fa.gpartf + j;
    double s = 0.0;
    int c = 0;
    for (; c + 8 <= fa.nchunks; c += 8) {
      double v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = ldcg(p + (int64_t)(c + k) * fa.npadf);
#pragma unroll
      for (int k = 0; k < 8; ++k) s += v[k];
    }
    for (; c < fa.nchunks; ++c) s += ldcg(p + (int64_t)c * fa.npadf);
    out[j] = s;
  }
}
// sum_c fpart[c], same bits in every thread of every CTA (lane-strided partial sums, xor butterfly, fixed order)
__device__ __forceinline__ double fused_fsum(const FusedArgs& fa, uint32_t scr) {
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < 32) {
    const double s = warp_sum(lane_strided_sum(fa.fpart, fa.nchunks, 1, lane));
    if (lane == 0) sts1(scr, s);
  }
  __syncthreads();
  const double r = lds1(scr);
  __syncthreads();
  return r;
}

// block / grid reductions for the 512-thread CTA (fixed order)
template <int K>
__device__ __forceinline__ void f_block_reduce_store(double (&v)[K], double* red, int G, int slot0, uint32_t scr) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const double s = warp_sum(v[k]);
    if (lane == 0) sts1(scr + (warp * K + k) * 8, s);
  }
  __syncthreads();
  if (threadIdx.x < K) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kFWarps; ++w) s += lds1(scr + (w * K + threadIdx.x) * 8);
    red[(int64_t)(slot0 + threadIdx.x) * G + blockIdx.x] = s;
  }
  __syncthreads();
}
template <int K>
__device__ __forceinline__ void f_grid_totals(const double* red, int G, int slot0, double (&out)[K], uint32_t scr) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = warp; k < K; k += kFWarps) {
    const double s = warp_sum(lane_strided_sum(red + (int64_t)(slot0 + k) * G, G, 1, lane));
    if (lane == 0) sts1(scr + k * 8, s);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) out[k] = lds1(scr + k * 8);
  __syncthreads();
}

// Grid-wide barrier on a monotonically increasing arrival counter.  The launch is sized by cudaOccupancyMaxActiveClusters
// (one CTA per SM) and, by default, carries the cooperative attribute next to the cluster dimension, so the runtime
// guarantees co-residency (api.cu: fused_config; ADAPROX_FUSED_NONCOOP=1 drops the attribute because Nsight Compute cannot
// replay a cooperative + cluster launch).  The row-sharded sweep-only launches are plain cluster launches.  In both cases
// the spin is BOUNDED: if the peers do not arrive within kGridBarTimeoutNs (another kernel holding SMs, MPS, a debugger),
// the CTA sets *err, stops waiting at every later barrier and the host reports ADAPROX_ERR_CUDA instead of hanging.
constexpr unsigned long long kGridBarTimeoutNs = 10000000000ull;
struct GridBar {
  unsigned long long* ctr;
  unsigned long long target;
  unsigned int G;
  int* err;
  __device__ __forceinline__ void sync() {
    __syncthreads();
    if (threadIdx.x == 0) {
      target += G;
      __threadfence();                                   // this CTA's global writes before its arrival
      atomicAdd(ctr, 1ull);
      unsigned long long seen, t0 = 0;
      for (unsigned spin = 0;; ++spin) {
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(ctr) : "memory");
        if (seen >= target) break;
        if ((spin & 4095u) == 4095u) {
          if (*reinterpret_cast<volatile int*>(err)) break;                    // somebody already gave up
          const unsigned long long now = globaltimer_ns();
          if (t0 == 0) t0 = now;
          else if (now - t0 > kGridBarTimeoutNs) { *reinterpret_cast<volatile int*>(err) = 1; __threadfence(); break; }
        }
      }
    }
    __syncthreads();
  }
};

// once per kernel, by all threads of every CTA: mbarriers, ragged-tail zero fill, cluster handshake
__device__ __forceinline__ void fused_smem_init(FusedSmem& fs, unsigned char* dyn_smem, unsigned char* s_ctl, int64_t ld) {
  fs.ring = smem_u32(dyn_smem);
  fs.ctl = smem_u32(s_ctl);
  fs.count = 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kFStages; ++s) { mbar_init(fs.ctl + kOffFull + 8 * s, 1); mbar_init(fs.ctl + kOffEmpty + 8 * s, kFXWarps); }
    for (int s = 0; s < kFDepth; ++s) mbar_init(fs.ctl + kOffCfull + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    // The last CTA of a cluster may own fewer than 8192 columns: the bulk copies never touch the tail of its ring
    // slots, so zero it once and the pass needs no column predicates (A = 0 there contributes nothing).
    int64_t width = ld - (int64_t)cluster_ctarank() * kFCols;
    width = width < 0 ? 0 : (width > kFCols ? kFCols : width);
    for (int s = 0; s < kFStages; ++s)
      for (int64_t j = width + threadIdx.x; j < kFCols; j += kFThreads) sts1(fs.ring + s * kFStageBytes + (uint32_t)j * 8, 0.0);
  }
  for (int j = threadIdx.x; j < (int)(kFDepth * kPartStride / 8); j += kFThreads) sts1(fs.ctl + kOffCpart + j * 8, 0.0);   // exchange buffers
  __syncthreads();
  cluster_arrive();          // every CTA of the cluster has initialised its shared memory before any peer writes into it
  cluster_wait();
}

// AdaPGM / fixed-step PGM (src/AdaProx.jl:312-364 with A = 0, h = Zero) around the fused pass.
__global__ void __launch_bounds__(kFThreads, 1) k_adapgm_fused(DProblem P, DOpts O, DWork W, FusedArgs fa) {
  GridBar grid{fa.bar, 0ull, gridDim.x, fa.err};
  const int b = blockIdx.x, G = gridDim.x;
  extern __shared__ __align__(1024) unsigned char dyn_smem[];
  __shared__ __align__(16) unsigned char s_ctl[kCtlBytes];
  __shared__ double s_scr[kFWarps * 8 + kMaxRed];
  FusedSmem fs;
  fused_smem_init(fs, dyn_smem, s_ctl, P.F.ld);
  const uint32_t scr = smem_u32(s_scr);

  const bool want_obj = O.want_objective != 0;
  int64_t j0, j1;
  cta_slice(P.n, b, G, j0, j1);
  unsigned long long sweep = 0;

  if (fa.sweep_only) {
    // Row-sharded solve: this launch only produces the shard's A'(Ax - b) partial and value sum for the all-reduce.
    // (A separate __global__ entry with the same body runs the sweep 1.7x slower for reasons not understood --
    // same SASS loops, same CTA placement, same data; see profiles/r01_notes.md -- so the sharded path shares
    // this entry point.)
    grid.target = fa.bar_base;
    const bool done = __ldcg(fa.sh_done) != 0;
    if (!done) {
      unsigned long long tq0 = 0;
      if (fa.lat && threadIdx.x == 0) tq0 = globaltimer_ns();
      fused_sweep(P.F, P.fvec, fa.sh_x, fs, fa, 0ull);
      if (fa.lat && threadIdx.x == 0) { fa.lat[4 * b + 1] = globaltimer_ns() - tq0; unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); fa.lat[4 * b + 3] = sm; }
    }
    grid.sync();
    if (fa.p2p.n <= 1) {
      if (!done) {
        fused_gradient_slice(fa, j0, j1, fa.sh_gbuf);
        const double f0 = fused_fsum(fa, scr);
        if (b == 0 && threadIdx.x == 0) { fa.sh_gbuf[P.n] = f0; fa.sh_gbuf[P.n + 1] = 0.0; }
      }
    } else {
      // All-reduce inside the kernel over NVLink peer memory (p2p.cuh).  Every rank holds the same `done`, and this
      // branch performs exactly three more grid barriers whether or not the solve has stopped (the host computes
      // fa.bar_base from a fixed number of barriers per launch).
      P2PState ps;
      p2p_begin(fa.p2p, ps);
      if (!done) {
        fused_gradient_slice(fa, j0, j1, fa.sh_gbuf);          // this rank's partial gradient (each CTA its own slice)
        const double f0 = fused_fsum(fa, scr);
        if (b == 0 && threadIdx.x == 0) { fa.sh_gbuf[P.n] = f0; fa.sh_gbuf[P.n + 1] = 0.0; }
      }
      grid.sync();                                             // the two value sums were written by CTA 0 only
      if (!done) p2p_allreduce<kFThreads>(fa.p2p, ps, grid, fa.sh_gbuf, fa.sh_gbuf, P.n + 2);
      else { grid.sync(); grid.sync(); }
    }
    cluster_arrive();          // no CTA exits while a peer could still address its shared memory
    cluster_wait();
    return;
  }

  double gamma, sigma, s0, s1;
  rule_init(O, gamma, sigma, s0, s1);
  int64_t n_eval = 0, n_grad = 0, n_proxg = 0, n_rec = 0;
  unsigned flags = 0;
  int xc = 0, gc = 0;
  double* x = W.xb[0];
  if (fa.hV > 0) {             // every CTA of this launch is running: the host may launch the helper CTAs now
    grid.sync();
    if (b == 0 && threadIdx.x == 0) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(fa.resident), "l"(fa.resident_seq) : "memory");
  }
  // Row-sharded solve, persistent form (fa.p2p.n > 1): this rank's matrix is the row block of a larger one.  After the sweep
  // the shard's partial gradient and value sum (n + 2 doubles) are all-reduced INSIDE this kernel over NVLink peer memory
  // (p2p.cuh); every rank obtains the same bits, so the replicated stepsize / prox arithmetic below stays in lock step and the
  // whole sharded solve is ONE launch per rank -- no NCCL, no host, no relaunch per iteration.
  const bool sharded = fa.p2p.n > 1;
  P2PState ps;
  p2p_begin(fa.p2p, ps);
  // dispenser / helper protocol (fa.tagged): CTA 0 arms the dispenser for sweep s BEFORE the grid barrier that precedes it and
  // raises `go` AFTER it (the iterate of sweep s is complete then); after a sweep every CTA waits until all chunks are done
  auto arm_sweep = [&](unsigned long long s_next) {
    if (fa.tagged && b == 0 && threadIdx.x == 0) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(fa.next), "l"(s_next << 32) : "memory");
  };
  auto release_sweep = [&](unsigned long long s_next) {
    if (fa.tagged && fa.hV > 0 && b == 0 && threadIdx.x == 0) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(fa.go), "l"(s_next) : "memory");
  };
  auto wait_chunks = [&](unsigned long long sweeps_done) {
    if (!(fa.tagged && fa.hV > 0)) return;                      // without helpers the grid barrier already covers every chunk
    if (threadIdx.x == 0) {
      const unsigned long long want = sweeps_done * (unsigned long long)fa.nchunks;
      unsigned long long seen, t0 = 0;
      for (unsigned spin = 0;; ++spin) {
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(fa.done) : "memory");
        if (seen >= want) break;
        if ((spin & 4095u) == 4095u) {
          if (*reinterpret_cast<volatile int*>(fa.err)) break;
          const unsigned long long now = globaltimer_ns();
          if (t0 == 0) t0 = now;
          else if (now - t0 > kGridBarTimeoutNs) { *reinterpret_cast<volatile int*>(fa.err) = 1; __threadfence(); break; }
        }
      }
    }
    __syncthreads();
  };
  // called after the sweep's grid barrier: gradient entries [j0, j1) -> out (this CTA's slice), returns sum of r_i^2 over ALL rows
  auto finish_gradient = [&](double* out) -> double {
    if (!sharded) {
      fused_gradient_slice(fa, j0, j1, out);
      return fused_fsum(fa, scr);
    }
    fused_gradient_slice(fa, j0, j1, fa.sh_gbuf);            // this rank's partial
    const double f0 = fused_fsum(fa, scr);
    if (b == 0 && threadIdx.x == 0) { fa.sh_gbuf[P.n] = f0; fa.sh_gbuf[P.n + 1] = 0.0; }
    grid.sync();                                             // the whole partial vector is in place
    p2p_allreduce<kFThreads>(fa.p2p, ps, grid, fa.sh_gbuf, fa.sh_gbuf, P.n + 2);
    for (int64_t j = j0 + threadIdx.x; j < j1; j += kFThreads) out[j] = ldcg(fa.sh_gbuf + j);
    return ldcg(fa.sh_gbuf + P.n);
  };

  // ---- prologue (:327-332) ----------------------------------------------------------------------------------
  {
    phase_stamp(W, 1, 1);
    fused_sweep(P.F, P.fvec, x, fs, fa, sweep++);        // sweep 0: dispenser armed and `go` = 0 by the host
  }
  grid.sync();
  wait_chunks(sweep);
  phase_stamp(W, 1, 2);
  {
    (void)finish_gradient(W.gb[gc]);
    double acc[1] = {0.0};
    double* xn = W.xb[1];
    for (int64_t j = j0 + threadIdx.x; j < j1; j += kFThreads) {
      const double vj = x[j] - gamma * W.gb[gc][j];                             // :330
      W.v[j] = vj;
      const double xj = prox_elem(P.g, vj, gamma, j, 0.0);                     // :332
      xn[j] = xj;
      if (want_obj) acc[0] += prox_value_elem(P.g, xj, j);
    }
    f_block_reduce_store<1>(acc, W.red, G, gval_slot(1), scr);
  }
  n_eval = 1; n_grad = 1; n_proxg = 1;
  arm_sweep(sweep);
  grid.sync();
  release_sweep(sweep);
  double* x_prev = W.xb[0];
  x = W.xb[1]; xc = 1;
  double* grad_prev = W.gb[0];
  double norm_res = INFINITY;
  int64_t it_done = O.maxit;
  bool converged = false;

  for (int64_t it = 1; it <= O.maxit; ++it) {
    if (p2p_failed(fa.p2p)) { flags |= ADAPROX_FLAG_COMM; it_done = it - 1; break; }   // a peer rank was lost (uniform: read after a grid barrier)
    phase_stamp(W, it, 0);
    {
      // :336 value + pullback in one sweep
      unsigned long long tq0 = 0;
      if (fa.lat && threadIdx.x == 0) tq0 = globaltimer_ns();
      fused_sweep(P.F, P.fvec, x, fs, fa, sweep++);
      if (fa.lat && threadIdx.x == 0) { fa.lat[4 * b + 1] = globaltimer_ns() - tq0; unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); fa.lat[4 * b + 3] = sm; }
    }
    n_eval++; n_grad++;
    grid.sync();
    wait_chunks(sweep);
    phase_stamp(W, it, 3);
    double* grad = W.gb[gc ^ 1];
    const double fsum_all = finish_gradient(grad);
    {
      double acc[4] = {0.0, 0.0, 0.0, 0.0};
      for (int64_t j = j0 + threadIdx.x; j < j1; j += kFThreads) {
        const double xj = x[j], gj = grad[j];
        const double pr = (W.v[j] - xj) / gamma + gj;                           // :338
        const double dg = gj - grad_prev[j], dx = xj - x_prev[j];
        acc[0] = fma(pr, pr, acc[0]);
        acc[1] = fma(dg, dg, acc[1]);
        acc[2] = fma(dg, dx, acc[2]);
        acc[3] = fma(dx, dx, acc[3]);
      }
      f_block_reduce_store<4>(acc, W.red, G, SLOT_PR, scr);
    }
    grid.sync();
    phase_stamp(W, it, 4);
    double t4[4], tg[1] = {0.0}, tf[1];
    f_grid_totals<4>(W.red, G, SLOT_PR, t4, scr);
    tf[0] = fsum_all;
    if (want_obj) f_grid_totals<1>(W.red, G, gval_slot(it), tg, scr);
    const double gamma_prev = gamma;
    rule_step(O, t4[1], t4[2], t4[3], gamma, sigma, s0, s1);                    // :341
    norm_res = sqrt(norm_sq_jl(t4[0]) + adapgm_dual_res_sq(gamma, gamma_prev, sigma));   // :348 (dual part: 0, or NaN -- phases.cuh)
    if (!(gamma == gamma) || !(norm_res == norm_res) || isinf(gamma)) flags |= ADAPROX_FLAG_NONFINITE;
    if (b == 0 && threadIdx.x == 0 && W.rec != nullptr && it <= O.max_records) {
      adaprox_record rc;
      rc.it = it; rc.gamma = gamma; rc.sigma = sigma; rc.norm_res = norm_res;
      rc.f_x = 0.5 * norm_sq_jl(tf[0]);
      rc.g_x = want_obj ? prox_value_finish(P.g.kind, P.g.lambda, tg[0]) : NAN;
      rc.h_Ax = want_obj ? 0.0 : NAN;
      rc.f_evals = n_eval; rc.grad_f_evals = n_grad; rc.prox_g_evals = n_proxg; rc.prox_h_evals = 0;
      rc.A_evals = 0; rc.At_evals = 0;
      W.rec[it - 1] = rc;
    }
    if (it <= O.max_records) n_rec = it;
    if (norm_res <= O.tol) { converged = true; it_done = it; break; }           // :354-356
    {
      double acc[1] = {0.0};
      double* xn = W.xb[(xc + 1) % 3];
      for (int64_t j = j0 + threadIdx.x; j < j1; j += kFThreads) {
        const double vj = x[j] - gamma * grad[j];                               // :359
        W.v[j] = vj;
        const double xj = prox_elem(P.g, vj, gamma, j, 0.0);                    // :361
        xn[j] = xj;
        if (want_obj) acc[0] += prox_value_elem(P.g, xj, j);
      }
      f_block_reduce_store<1>(acc, W.red, G, gval_slot(it + 1), scr);
    }
    n_proxg++;
    x_prev = x; xc = (xc + 1) % 3; x = W.xb[xc];
    grad_prev = grad; gc ^= 1;
    arm_sweep(sweep);
    grid.sync();
    release_sweep(sweep);
    phase_stamp(W, it, 7);
  }
  // helpers leave when the dispenser carries the exit tag
  if (fa.tagged && b == 0 && threadIdx.x == 0) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(fa.next), "l"(0xffffffffull << 32) : "memory");
    if (fa.hV > 0) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(fa.go), "l"(~0ull) : "memory");
  }

  const int64_t tid = (int64_t)b * kFThreads + threadIdx.x, nt = (int64_t)G * kFThreads;
  for (int64_t j = tid; j < P.n; j += nt) W.xout[j] = x[j];
  if (b == 0 && threadIdx.x == 0) {
    DResult r;
    r.iters = it_done;
    r.flags = flags | (converged ? ADAPROX_FLAG_CONVERGED : 0u);
    r.xbuf = 0;
    r.f_evals = n_eval; r.grad_f_evals = n_grad; r.prox_g_evals = n_proxg; r.prox_h_evals = 0;
    r.A_evals = 0; r.At_evals = 0; r.n_records = n_rec;
    r.final_gamma = gamma; r.final_sigma = sigma; r.final_norm_res = norm_res;
    *W.res = r;
  }
  cluster_arrive();          // no CTA exits while a peer could still address its shared memory
  cluster_wait();
}

}  // namespace adaprox

#include "solver_fused_helper.cuh"
