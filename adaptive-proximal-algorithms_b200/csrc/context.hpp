// context.hpp -- host-side state behind an adaprox_handle.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <vector>
#include "phases.cuh"

namespace adaprox {

struct HostMatrix {
  DMat d{};
  std::vector<void*> allocs;     // device allocations owned by this matrix
  int64_t m_global = 0, row0 = 0;
  bool sharded = false;
};

struct HostVector {
  double* p = nullptr;
  int64_t len = 0;
};

struct Comm;   // comm.cu

}  // namespace adaprox

struct adaprox_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;    // helper CTAs of the single-sweep kernel run beside it (solver_fused_helper.cuh)
  cudaEvent_t ev_h = nullptr;
  unsigned long long* resident_host = nullptr;   // pinned, mapped: the cluster kernel reports "all my CTAs are running" (helper launch gate)
  unsigned long long* resident_dev = nullptr;
  unsigned long long solve_seq = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int sm_count = 0, cc_major = 0, cc_minor = 0;
  int grid = 0;                  // CTAs of every persistent kernel (SMs x resident CTAs)
  std::string err;
  std::map<int64_t, adaprox::HostMatrix> mats;
  std::map<int64_t, adaprox::HostVector> vecs;
  int64_t next_id = 1;
  // grow-only workspace arena
  char* ws = nullptr;
  size_t ws_bytes = 0, ws_used = 0;
  int64_t launches = 0;
  adaprox::Comm* comm = nullptr;
};

namespace adaprox {

inline int fail(adaprox_ctx* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}

#define AP_CUDA(h, call)                                                                         \
  do {                                                                                           \
    cudaError_t e__ = (call);                                                                    \
    if (e__ != cudaSuccess)                                                                      \
      return adaprox::fail(h, ADAPROX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
  } while (0)

}  // namespace adaprox
