// p2p_args.cuh -- kernel-side view of the peer-mapped exchange blocks (see p2p.cuh for the protocol)
#pragma once
#include <stdint.h>

namespace adaprox {

constexpr int kP2PMaxRanks = 8;

struct P2PArgs {
  int n, rank;                                   // n <= 1: not sharded / not attached
  int64_t cap;
  double* buf[kP2PMaxRanks][2];                  // rank q's two exchange buffers as mapped here
  unsigned long long* flags[kP2PMaxRanks];       // rank q's flag words as mapped here
  int* err;
};

}  // namespace adaprox
