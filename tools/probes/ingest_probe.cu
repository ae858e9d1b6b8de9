// ingest_probe.cu -- how fast can G SMs (one CTA each) pull an HBM-resident stream into shared memory with cp.async.bulk,
// as a function of the number of SMs used, the tile size and the ring depth?  Stand-alone (NOT part of the library).
// The single-sweep kernel (solver_fused.cuh) runs 112 of 148 SMs with a 3 x 64 KB ring; this probe separates "the SM's
// ingest port is the limit" from "the ring is too shallow".  Consumer work is one LDS.128 per 16 bytes (as the dot
// warps of the real kernel), no arithmetic worth mentioning.
// Build:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o ingest_probe ingest_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e__)); std::exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// CTA b streams tiles b, b + G, b + 2G, ... of `tile` bytes each
__global__ void __launch_bounds__(256, 1) k_stream(const char* __restrict__ src, long long ntiles, int tile, int stages, int touch, double* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ unsigned long long bars[32];
  const uint32_t ring = smem_u32(smem), full = smem_u32(bars);
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full + 8 * s), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long mine = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  auto issue = [&](long long t) {
    const int s = (int)(t % stages);
    const char* p = src + (size_t)(blockIdx.x + t * gridDim.x) * (size_t)tile;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full + 8 * s), "r"(tile) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(ring + s * tile), "l"(p), "r"(tile), "r"(full + 8 * s) : "memory");
  };
  double acc = 0.0;
  if (threadIdx.x == 0) for (long long t = 0; t < stages - 1 && t < mine; ++t) issue(t);
  for (long long t = 0; t < mine; ++t) {
    if (threadIdx.x == 0 && t + stages - 1 < mine) issue(t + stages - 1);
    const int s = (int)(t % stages);
    const uint32_t ph = (uint32_t)((t / stages) & 1);
    uint32_t ok;
    do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(full + 8 * s), "r"(ph) : "memory");
    } while (!ok);
    for (int rep = 0; rep < touch; ++rep)            // touch = how many times every byte of the tile is read back with LDS.128
      for (int k = 0; k < tile / (256 * 16); ++k) {
        double2 v;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(ring + s * tile + (k * 256 + threadIdx.x) * 16));
        acc += v.x + v.y;
      }
    __syncthreads();
  }
  if (acc == 123.456) sink[0] = acc;
}

// steady mode: ./ingest_probe steady <ctas> <tile_KB> <stages> <lds passes> <seconds>: one configuration back to back (power measurements)
int main(int argc, char** argv) {
  if (argc >= 7 && std::string(argv[1]) == "steady") {
    const int G = std::atoi(argv[2]), tile = std::atoi(argv[3]) * 1024, stages = std::atoi(argv[4]), touch = std::atoi(argv[5]);
    const double secs = std::atof(argv[6]);
    const size_t bytes = (size_t)16 << 30;
    char* d_src; double* d_sink;
    CK(cudaMalloc(&d_src, bytes)); CK(cudaMemset(d_src, 0, bytes)); CK(cudaMalloc(&d_sink, 8));
    CK(cudaFuncSetAttribute((const void*)k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int reps = (int)(secs / 0.0025) + 1;
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) k_stream<<<G, 256, stages * tile>>>(d_src, (long long)(bytes / tile), tile, stages, touch, d_sink);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    std::printf("{\"probe\": \"steady HBM->smem stream\", \"ctas\": %d, \"tile_KB\": %d, \"stages\": %d, \"lds_reads_per_byte\": %d, \"seconds\": %.2f, \"GBps\": %.0f}\n",
                G, tile / 1024, stages, touch, ms * 1e-3, (double)bytes * reps / (ms * 1e-3) / 1e9);
    return 0;
  }
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  const size_t bytes = (size_t)16 << 30;                 // 16 GB stream: far larger than L2
  char* d_src; double* d_sink;
  CK(cudaMalloc(&d_src, bytes)); CK(cudaMemset(d_src, 0, bytes)); CK(cudaMalloc(&d_sink, 8));
  CK(cudaFuncSetAttribute((const void*)k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  std::printf("{\"device\": \"%s\", \"sms\": %d, \"stream_GB\": %.1f}\n", prop.name, sms, bytes / 1e9);
  const int grids[] = {32, 64, 96, 112, 128, sms};
  struct Ring { int tile_kb, stages; };
  const Ring rings[] = {{64, 3}, {64, 2}, {32, 6}, {32, 4}};
  for (int touch : {0, 1, 2, 3})
  for (int G : grids)
    for (const Ring& r : rings) {
      const int tile = r.tile_kb * 1024;
      const long long ntiles = (long long)(bytes / tile);
      float best = 1e30f;
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0));
        k_stream<<<G, 256, r.stages * tile>>>(d_src, ntiles, tile, r.stages, touch, d_sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        best = ms < best ? ms : best;
      }
      CK(cudaGetLastError());
      std::printf("{\"probe\": \"HBM->smem bulk-copy stream\", \"lds_reads_per_byte\": %d, \"ctas\": %d, \"tile_KB\": %d, \"stages\": %d, \"in_flight_KB\": %d, \"ms\": %.2f, \"GBps\": %.0f, \"GBps_per_sm\": %.1f}\n",
                  touch, G, r.tile_kb, r.stages, (r.stages - 1) * r.tile_kb, best, bytes / (best * 1e-3) / 1e9, bytes / (best * 1e-3) / 1e9 / G);
      std::fflush(stdout);
    }
  return 0;
}
