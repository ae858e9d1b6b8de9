"""Row-sharded fused solve on ONE GPU (communicator of size 1) -- isolates the per-launch kernel from the collective."""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT="29655")
import torch.distributed as dist  # noqa: E402
import adaprox_b200 as AdaProx  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dist.init_process_group("gloo", rank=0, world_size=1)
dev = AdaProx.Device(0)
AdaProx.set_default_device(dev)
AdaProx.sharding.attach_communicator(dev, dist)
P = AdaProx.generate_planted_lasso(65536, 131072, pfactor=5, seed=0, power_iters=2, row0=0, rows=rows, dev=dev)
f, g = AdaProx.LinearLeastSquares(P["A"], P["b"]), AdaProx.NormL1(1.0)
for k in range(2):
    x, it = AdaProx.adaptive_proxgrad(np.zeros(131072), f=f, g=g, rule=AdaProx.OurRule(gamma=1 / P["Lf"]), tol=0.0, maxit=steps)
    info = AdaProx.last_solve_info()
    print("rows", rows, "ms/iter", info["solve_ms"] / steps, info)
dist.destroy_process_group()
