// fp64_peak_probe.cu -- stand-alone measurement of the fp64 pipes of one B200 (NOT part of libadaprox_cuda.so):
//   1. DFMA (vector pipe): 8 independent chains per thread, 1024 threads per SM
//   2. DMMA  mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4, the only fp64 MMA shape sm_100a has: the PTX shapes
//      m16n8k4 / k8 / k16 compile to runs of DMMA.8x8x4), 4 / 8 / 16 warps per SM, 4 independent accumulators
//   3. DFMA and DMMA issued from different warps of the same SM (do the pipes add up?)
//   4. resident clusters (cudaOccupancyMaxActiveClusters) for cluster sizes 1..16 of a 512-thread, 200 KB CTA
// Output: one JSON line per measurement.  `profiles/r02_fp64_peaks.json` is this program's output on the pool's B200.
// Build:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o fp64_peak_probe fp64_peak_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e__)); std::exit(1); } } while (0)

__global__ void __launch_bounds__(1024, 1) k_dfma(int reps, double seed, double* sink) {
  double a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = seed + threadIdx.x * 1e-9 + k;
  const double m = 1.0 + seed * 1e-12, c = seed * 1e-13;
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] = fma(a[k], m, c);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += a[k];
  if (s == 123.456) sink[0] = s;
}

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// mode 0: every warp DMMA;  mode 1: even warps DMMA, odd warps DFMA
__global__ void __launch_bounds__(1024, 1) k_dmma(int reps, int mode, double seed, double* sink) {
  const int warp = threadIdx.x >> 5;
  double s = 0.0;
  if (mode == 0 || (warp & 1) == 0) {
    double c[8][2];
#pragma unroll
    for (int k = 0; k < 8; ++k) { c[k][0] = 0.0; c[k][1] = 0.0; }
    const double a = seed + threadIdx.x * 1e-9, b = 1.0 + seed * 1e-12;
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int k = 0; k < 8; ++k) dmma884(c[k], a, b);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) s += c[k][0] + c[k][1];
  } else {
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = seed + threadIdx.x * 1e-9 + k;
    const double m = 1.0 + seed * 1e-12, c = seed * 1e-13;
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = fma(a[k], m, c);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
  }
  if (s == 123.456) sink[0] = s;
}

__global__ void __launch_bounds__(512, 1) k_dummy(double* sink) {
  extern __shared__ double sm[];
  if (threadIdx.x == 9999) sink[0] = sm[0];
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  int clk_khz = 0;
  CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  std::printf("{\"device\": \"%s\", \"sms\": %d, \"max_sm_clock_mhz\": %.0f}\n", prop.name, sms, clk_khz * 1e-3);
  double* d_sink;
  CK(cudaMalloc(&d_sink, 8));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto timed = [&](auto launch) {
    launch();                                     // warm-up
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int t = 0; t < 5; ++t) {
      CK(cudaEventRecord(e0));
      launch();
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      best = ms < best ? ms : best;
    }
    CK(cudaGetLastError());
    return (double)best;
  };
  {
    const int reps = 20000;
    for (int threads : {256, 512, 1024}) {
      const double ms = timed([&] { k_dfma<<<sms, threads>>>(reps, 1.0, d_sink); });
      const double flop = 2.0 * 64.0 * reps * threads * (double)sms;
      std::printf("{\"probe\": \"DFMA vector pipe\", \"threads_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", threads, ms, flop / (ms * 1e-3) / 1e12);
    }
  }
  {
    const int reps = 20000;
    for (int threads : {128, 256, 512, 1024}) {
      const double ms = timed([&] { k_dmma<<<sms, threads>>>(reps, 0, 1.0, d_sink); });
      const double flop = 2.0 * 8 * 8 * 4 * 32.0 * reps * (threads / 32) * (double)sms;
      std::printf("{\"probe\": \"DMMA.8x8x4 tensor pipe\", \"warps_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", threads / 32, ms, flop / (ms * 1e-3) / 1e12);
    }
    for (int threads : {256, 512, 1024}) {
      const double ms = timed([&] { k_dmma<<<sms, threads>>>(reps, 1, 1.0, d_sink); });
      const double w = threads / 32 / 2;
      const double flop_mma = 2.0 * 8 * 8 * 4 * 32.0 * reps * w * (double)sms;
      const double flop_fma = 2.0 * 32.0 * reps * (w * 32) * (double)sms;
      std::printf("{\"probe\": \"DMMA + DFMA in different warps\", \"warps_per_sm\": %d, \"ms\": %.3f, \"tflops_dmma\": %.2f, \"tflops_dfma\": %.2f, \"tflops_sum\": %.2f}\n",
                  threads / 32, ms, flop_mma / (ms * 1e-3) / 1e12, flop_fma / (ms * 1e-3) / 1e12, (flop_mma + flop_fma) / (ms * 1e-3) / 1e12);
    }
  }
  {
    const int smem = 200 * 1024;
    CK(cudaFuncSetAttribute((const void*)k_dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute((const void*)k_dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    for (int C : {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 16}) {
      cudaLaunchConfig_t cfg = {};
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.gridDim = dim3(C); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem; cfg.attrs = at; cfg.numAttrs = 1;
      int q = 0;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&q, (const void*)k_dummy, &cfg);
      if (e != cudaSuccess) { std::printf("{\"probe\": \"resident clusters\", \"cluster\": %d, \"error\": \"%s\"}\n", C, cudaGetErrorString(e)); cudaGetLastError(); continue; }
      std::printf("{\"probe\": \"resident clusters (512 threads, 200 KB smem)\", \"cluster\": %d, \"clusters\": %d, \"sms_used\": %d}\n", C, q, q * C);
    }
  }
  return 0;
}
