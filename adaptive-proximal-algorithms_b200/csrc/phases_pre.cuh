// phases_pre.cuh -- the part of a smooth-term evaluation that has to run BEFORE phase A.
//
// Only the Gram form of the Quadratic term needs one (ADAPROX_F_QUADRATIC_GRAM): the dual SVM of
// dual_svm/runme.jl:47-49 builds Q = Dy*X*X'*Dy = Z*Z' with Z = Dy*X (N x d) and then multiplies by the N x N matrix
// (8 N^2 bytes per evaluation).  Q*x = Z*(Z'*x) needs two sweeps over Z instead (16 N d bytes: 12.5x fewer at
// 50000 x 2000):
//   pre-1   gpart = partials of Z' * x[rows of this rank]                 (gemv_t_phase)
//   pre-2   u = sum of the partials (d entries, fixed order); row-sharded Z: u summed over the ranks inside the
//           kernel (p2p.cuh), the same bits on every rank
//   A       zpart = partials of Z * u                                     (gemv_n_phase, in f_phase_A)
//   B       temp_i = (Z u)_i, value sums, exactly as for the dense-Q form  (f_phase_B)
// Every other f kind: no-op (uniform branch, no barrier).
#pragma once
#include "p2p.cuh"

namespace adaprox {

// `x` is the full-length (replicated) primal vector.  Contains 2 grid barriers (+ 2 inside the all-reduce when sharded).
// `ps` may be null when the caller never runs sharded (operator kernel, malitsky_pock).
template <class Grid>
__device__ __forceinline__ void f_phase_pre(Grid& grid, const DProblem& P, const DWork& W, const double* x, Sh& sh, int b,
                                            int G, P2PState* ps) {
  if (P.f_kind != ADAPROX_F_QUADRATIC_GRAM) return;
  gemv_t_phase(P.F, x + P.f_row0, sh, b, G);
  grid.sync();
  int64_t k0, k1;
  cta_slice(P.F.n, b, G, k0, k1);
  gsum_slice(P.F, k0, k1, W.fu, G);
  if (ps != nullptr && P.p2p.n > 1 && P.F_sharded) p2p_allreduce<kThreads>(P.p2p, *ps, grid, W.fu, W.fu, P.F.n);
  else grid.sync();
}

}  // namespace adaprox
