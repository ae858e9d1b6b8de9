// p2p.cuh -- all-reduce INSIDE a grid-synchronous kernel over NVLink peer memory (no NCCL, no host).
//
// Every rank owns an exchange block (cudaMalloc + CUDA IPC, mapped by all peers; comm.inl):
//   u64 words:  [q]      "rank q has published exchange T"        (value T + 1, written by rank q into every block)
//               [8 + q]  "rank q has finished READING exchange T" (value T + 1)
//               [16]     number of exchanges this rank has completed (read once at kernel start)
//   then two buffers of `cap` doubles (exchange T uses buffer T & 1).
// All ranks run the same kernels on replicated control state, so they perform the same sequence of exchanges; T is
// carried in registers (P2PState) and is identical everywhere.  Sums are formed in rank order: the same bits on
// every rank.  Buffer reuse is safe because a rank overwrites buffer T & 1 only after every peer reported that it
// finished reading exchange T - 1 (hence T - 2).  Spins are bounded (5 s): on timeout *err is set, later waits return at
// once, the persistent kernels leave their loop (p2p_failed) and the host returns ADAPROX_ERR_COMM; adaprox_p2p_reset
// (every rank, then a host barrier) makes the blocks usable again.
#pragma once
#include "phases.cuh"      // pulls in p2p_args.cuh

namespace adaprox {

struct P2PState { unsigned long long T; };

__device__ __forceinline__ void p2p_begin(const P2PArgs& pa, P2PState& st) {
  st.T = 0;
  if (pa.n > 1) asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(st.T) : "l"(pa.flags[pa.rank] + 16) : "memory");
}
// Has an exchange of this solve timed out (on this rank)?  Uniform over the grid when read after the last grid barrier of
// an exchange: the persistent kernels test it at the top of every iteration and leave the loop with ADAPROX_FLAG_COMM.
__device__ __forceinline__ bool p2p_failed(const P2PArgs& pa) { return pa.n > 1 && *reinterpret_cast<volatile int*>(pa.err) != 0; }
__device__ __forceinline__ void p2p_wait_ge(const P2PArgs& pa, const unsigned long long* src, unsigned long long v) {
  if (*reinterpret_cast<volatile int*>(pa.err) != 0) return;       // a peer was lost earlier: never wait again (5 s each otherwise)
  const unsigned long long t0 = globaltimer_ns();
  unsigned long long seen;
  for (unsigned spin = 0;; ++spin) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(src) : "memory");
    if (seen >= v) break;
    if ((spin & 1023u) == 1023u && globaltimer_ns() - t0 > 5000000000ull) { *reinterpret_cast<volatile int*>(pa.err) = 1; __threadfence(); break; }
  }
}
__device__ __forceinline__ double p2p_ld(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void p2p_publish(const P2PArgs& pa, int word0, unsigned long long want) {
  // one thread per peer (threads 0 .. n-1 of CTA 0)
  __threadfence_system();
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(pa.flags[threadIdx.x] + word0 + pa.rank), "l"(want) : "memory");
}

// vec[0 .. count) <- sum over ranks, in place.  Called by every thread of every CTA; `src` may differ from `vec`
// (src = this rank's partial, written by CTA b for its cta_slice of [0, count) or made visible by an earlier grid
// barrier).  NT = threads per CTA.  Contains two grid barriers.
template <int NT, class Grid>
__device__ __forceinline__ void p2p_allreduce(const P2PArgs& pa, P2PState& st, Grid& grid, const double* src, double* vec, int64_t count) {
  const int b = blockIdx.x, G = gridDim.x;
  const unsigned long long T = st.T, want = T + 1;
  unsigned long long* myflags = pa.flags[pa.rank];
  double* mine = pa.buf[pa.rank][T & 1];
  int64_t j0, j1;
  cta_slice(count, b, G, j0, j1);
  if (threadIdx.x < pa.n) p2p_wait_ge(pa, myflags + 8 + threadIdx.x, T);       // peers finished reading exchange T - 1
  __syncthreads();
  for (int64_t j = j0 + threadIdx.x; j < j1; j += NT) mine[j] = src[j];
  grid.sync();
  if (b == 0 && threadIdx.x < pa.n) p2p_publish(pa, 0, want);
  if (threadIdx.x < pa.n) p2p_wait_ge(pa, myflags + threadIdx.x, want);
  __syncthreads();
  for (int64_t j = j0 + threadIdx.x; j < j1; j += NT) {
    double s = 0.0;
    for (int q = 0; q < pa.n; ++q) s += p2p_ld(pa.buf[q][T & 1] + j);
    vec[j] = s;
  }
  grid.sync();
  if (b == 0 && threadIdx.x < pa.n) {
    p2p_publish(pa, 8, want);
    if (threadIdx.x == 0) asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(myflags + 16), "l"(want) : "memory");
  }
  st.T = want;
}

// K scalars that every thread of every CTA holds with identical bits -> their sums over the ranks (rank order), again
// identical in every thread of every rank.  Two grid barriers.
template <int K, class Grid>
__device__ __forceinline__ void p2p_allreduce_scalars(const P2PArgs& pa, P2PState& st, Grid& grid, double (&vals)[K]) {
  const int b = blockIdx.x;
  const unsigned long long T = st.T, want = T + 1;
  unsigned long long* myflags = pa.flags[pa.rank];
  double* mine = pa.buf[pa.rank][T & 1];
  if (threadIdx.x < pa.n) p2p_wait_ge(pa, myflags + 8 + threadIdx.x, T);
  __syncthreads();
  if (b == 0 && threadIdx.x < K) mine[threadIdx.x] = vals[threadIdx.x];
  grid.sync();
  if (b == 0 && threadIdx.x < pa.n) p2p_publish(pa, 0, want);
  if (threadIdx.x < pa.n) p2p_wait_ge(pa, myflags + threadIdx.x, want);
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double s = 0.0;
    for (int q = 0; q < pa.n; ++q) s += p2p_ld(pa.buf[q][T & 1] + k);
    vals[k] = s;
  }
  grid.sync();
  if (b == 0 && threadIdx.x < pa.n) {
    p2p_publish(pa, 8, want);
    if (threadIdx.x == 0) asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(myflags + 16), "l"(want) : "memory");
  }
  st.T = want;
}

}  // namespace adaprox
