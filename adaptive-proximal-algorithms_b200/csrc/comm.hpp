// comm.hpp -- row-sharded multi-GPU support: NCCL (loaded lazily with dlopen so a
// single-GPU user needs no NCCL at all) and the split-phase sharded solver.
#pragma once
#include "context.hpp"

namespace adaprox {

int comm_allreduce_sum(adaprox_ctx* h, double* buf_dev, int64_t count);
void comm_destroy(adaprox_ctx* h);
struct P2PArgs;
void p2p_fill(adaprox_ctx* h, P2PArgs* pa);            // kernel-side view of the peer-mapped exchange blocks (p2p.cuh)
bool p2p_ready(adaprox_ctx* h, int64_t count);         // attached and large enough for vectors of `count` doubles
int p2p_check(adaprox_ctx* h);                         // error flag of the bounded spins
int comm_nranks(adaprox_ctx* h);                       // ranks of the attached communicator (1 without one)
int comm_setup_kernels();      // opt the split-phase kernels into the ring's dynamic shared memory

int solve_sharded(adaprox_ctx* h, const adaprox_problem* p, const adaprox_options* o, const DProblem& P, const DOpts& O,
                  HostMatrix* fm, HostMatrix* am, const double* x0, const double* y0, double* x_out, double* y_out,
                  adaprox_record* records, adaprox_result* res);

}  // namespace adaprox
