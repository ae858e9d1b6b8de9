// latency_probe.cu -- stand-alone micro-benchmarks behind DESIGN.md section 10 (NOT part of libadaprox_cuda.so):
//   1. cooperative-groups grid.sync() round trip for G = 16 ... 2 x SMs CTAs of 256 threads
//   2. a hand-rolled arrival-counter barrier (the GridBar of solver_fused.cuh) on the same grids
//   3. hardware cluster barrier (barrier.cluster.arrive.release / wait.acquire) for cluster sizes 2, 4, 8, 16
//   4. L2 -> shared-memory ingest of ONE CTA with cp.async.bulk as a function of the bytes kept in flight
//      (1 ... 12 stages of 16 KB) on an L2-resident 4 MB buffer: how deep a ring a small-problem kernel needs
//   5. DSMEM: st.async to a peer CTA + remote mbarrier completion, round trip between two CTAs of a cluster
// Build:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o latency_probe latency_probe.cu
// Run:    ./latency_probe        (prints one line per measurement; a few hundred ms in total)
// Written at the end of round 1 without a GPU at hand (compiles, SASS checked); first numbers belong to round 2.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e__)); std::exit(1); } } while (0)

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- 1. grid.sync ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) k_gridsync(int reps, unsigned long long* out) {
  cg::grid_group grid = cg::this_grid();
  grid.sync();
  const unsigned long long t0 = gtime();
  for (int r = 0; r < reps; ++r) grid.sync();
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = gtime() - t0;
}

// ---- 2. arrival counter ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) k_counterbar(int reps, unsigned long long* ctr, unsigned long long* out) {
  cg::grid_group grid = cg::this_grid();      // cooperative launch only to guarantee co-residency
  const unsigned long long G = gridDim.x;
  unsigned long long target = 0;
  auto bar = [&]() {
    __syncthreads();
    if (threadIdx.x == 0) {
      target += G;
      __threadfence();
      atomicAdd(ctr, 1ull);
      unsigned long long seen;
      do { asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(ctr) : "memory"); } while (seen < target);
    }
    __syncthreads();
  };
  bar();
  const unsigned long long t0 = gtime();
  for (int r = 0; r < reps; ++r) bar();
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = gtime() - t0;
  (void)grid;
}

// ---- 3. cluster barrier ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1) k_clusterbar(int reps, unsigned long long* out) {
  auto bar = []() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  };
  bar();
  const unsigned long long t0 = gtime();
  for (int r = 0; r < reps; ++r) bar();
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = gtime() - t0;
}

// ---- 4. one CTA streaming an L2-resident buffer through a ring of `stages` x 16 KB ------------------------------------
constexpr int kTile = 16384;
__global__ void __launch_bounds__(256, 1) k_ingest(const double* __restrict__ src, long long tiles_total, int tiles_in_buf, int stages,
                                                   unsigned long long* out, double* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ unsigned long long bars[16];
  const uint32_t ring = smem_u32(smem), full = smem_u32(bars);
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full + 8 * s), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](long long t) {
    const int s = (int)(t % stages);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full + 8 * s), "r"(kTile) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(ring + s * kTile), "l"(src + (t % tiles_in_buf) * (kTile / 8)), "r"(kTile), "r"(full + 8 * s) : "memory");
  };
  double acc = 0.0;
  const unsigned long long t0 = gtime();
  if (threadIdx.x == 0) for (long long t = 0; t < stages - 1 && t < tiles_total; ++t) issue(t);
  for (long long t = 0; t < tiles_total; ++t) {
    if (threadIdx.x == 0 && t + stages - 1 < tiles_total) issue(t + stages - 1);
    const int s = (int)(t % stages);
    const uint32_t ph = (uint32_t)((t / stages) & 1);
    uint32_t ok;
    do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(full + 8 * s), "r"(ph) : "memory");
    } while (!ok);
    // touch the tile (one LDS.128 per thread per 4 KB) so that the consumer side is not free
    for (int k = 0; k < kTile / (256 * 16); ++k) {
      double2 v;
      asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(ring + s * kTile + (k * 256 + threadIdx.x) * 16));
      acc += v.x + v.y;
    }
    __syncthreads();               // slot s is free again (the next issue into it happens after this point)
  }
  if (threadIdx.x == 0) out[0] = gtime() - t0;
  if (acc == 123.456) sink[0] = acc;
}

// ---- 5. DSMEM ping-pong between CTA 0 and CTA 1 of a cluster ----------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(32, 1) k_dsmem_pingpong(int reps, unsigned long long* out) {
  __shared__ __align__(8) unsigned long long bar;
  __shared__ __align__(8) double slot;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const uint32_t bar_a = smem_u32(&bar), slot_a = smem_u32(&slot);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_a), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  uint32_t peer_bar, peer_slot;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer_bar) : "r"(bar_a), "r"(rank ^ 1u));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer_slot) : "r"(slot_a), "r"(rank ^ 1u));
  unsigned long long t0 = 0;
  if (threadIdx.x == 0) {
    t0 = gtime();
    for (int r = 0; r < reps; ++r) {
      const uint32_t ph = (uint32_t)(r & 1);
      if (rank == 0) {            // send, then wait for the echo
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 8;" ::"r"(bar_a) : "memory");
        asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" ::"r"(peer_slot), "d"((double)r), "r"(peer_bar) : "memory");
      } else {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 8;" ::"r"(bar_a) : "memory");
      }
      uint32_t ok;
      do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar_a), "r"(ph) : "memory");
      } while (!ok);
      if (rank == 1)              // echo
        asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" ::"r"(peer_slot), "d"((double)r), "r"(peer_bar) : "memory");
    }
    if (rank == 0) out[0] = gtime() - t0;
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  unsigned long long *d_out, *d_ctr, h_out = 0;
  CK(cudaMalloc(&d_out, 8)); CK(cudaMalloc(&d_ctr, 8));
  const int reps = 2000;
  std::printf("{\"device\": \"%s\", \"sms\": %d}\n", prop.name, sms);

  const int grids[] = {16, 37, 74, 111, sms, 2 * sms};
  for (int G : grids) {
    int r = reps; void* a1[] = {&r, &d_out};
    CK(cudaLaunchCooperativeKernel((const void*)k_gridsync, dim3(G), dim3(256), a1, 0, 0));
    CK(cudaMemcpy(&h_out, d_out, 8, cudaMemcpyDeviceToHost));
    std::printf("{\"probe\": \"cg grid.sync\", \"ctas\": %d, \"ns_per_barrier\": %.1f}\n", G, (double)h_out / reps);
    CK(cudaMemset(d_ctr, 0, 8));
    void* a2[] = {&r, &d_ctr, &d_out};
    CK(cudaLaunchCooperativeKernel((const void*)k_counterbar, dim3(G), dim3(256), a2, 0, 0));
    CK(cudaMemcpy(&h_out, d_out, 8, cudaMemcpyDeviceToHost));
    std::printf("{\"probe\": \"arrival-counter barrier\", \"ctas\": %d, \"ns_per_barrier\": %.1f}\n", G, (double)h_out / reps);
  }

  CK(cudaFuncSetAttribute((const void*)k_clusterbar, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  for (int C : {2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3(C); cfg.blockDim = dim3(256); cfg.attrs = at; cfg.numAttrs = 1;
    int r = reps; void* a3[] = {&r, &d_out};
    cudaError_t e = cudaLaunchKernelExC(&cfg, (const void*)k_clusterbar, a3);
    if (e != cudaSuccess) { std::printf("{\"probe\": \"cluster barrier\", \"cluster\": %d, \"error\": \"%s\"}\n", C, cudaGetErrorString(e)); cudaGetLastError(); continue; }
    CK(cudaMemcpy(&h_out, d_out, 8, cudaMemcpyDeviceToHost));
    std::printf("{\"probe\": \"cluster barrier\", \"cluster\": %d, \"ns_per_barrier\": %.1f}\n", C, (double)h_out / reps);
  }

  {
    const int tiles_in_buf = 256;                       // 4 MB: L2-resident after the first sweep
    const long long tiles_total = 4096;                 // 64 MB streamed
    double *d_src, *d_sink;
    CK(cudaMalloc(&d_src, (size_t)tiles_in_buf * kTile)); CK(cudaMemset(d_src, 0, (size_t)tiles_in_buf * kTile)); CK(cudaMalloc(&d_sink, 8));
    CK(cudaFuncSetAttribute((const void*)k_ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * kTile));
    for (int stages : {2, 3, 4, 6, 8, 12}) {
      for (int pass = 0; pass < 2; ++pass) {            // pass 0 warms L2
        k_ingest<<<1, 256, stages * kTile>>>(d_src, tiles_total, tiles_in_buf, stages, d_out, d_sink);
        CK(cudaDeviceSynchronize());
      }
      CK(cudaMemcpy(&h_out, d_out, 8, cudaMemcpyDeviceToHost));
      std::printf("{\"probe\": \"one-CTA L2->smem ingest\", \"stages_16KB\": %d, \"GB_per_s\": %.1f, \"us_per_tile\": %.3f}\n", stages,
                  (double)tiles_total * kTile / (double)h_out, (double)h_out * 1e-3 / tiles_total);
    }
  }

  {
    int r = reps; 
    k_dsmem_pingpong<<<2, 32>>>(r, d_out);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h_out, d_out, 8, cudaMemcpyDeviceToHost));
    std::printf("{\"probe\": \"DSMEM st.async ping-pong (2 CTAs)\", \"ns_round_trip\": %.1f}\n", (double)h_out / reps);
  }
  return 0;
}
