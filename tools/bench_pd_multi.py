"""BASELINE configs[2] (LAD, 50000 x 2001, AdaPDM+) with the data matrix row-sharded over N GPUs; the persistent kernel
all-reduces A'y and the dual sums over NVLink peer memory itself (no NCCL, no host in the loop).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_pd_multi.py"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import adaprox_b200 as AdaProx  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group(backend="gloo")
    dev = AdaProx.Device(local)
    AdaProx.set_default_device(dev)
    m, d = int(os.environ.get("PD_ROWS", "50000")), 2000
    X, yv = AdaProx.synth.dense_regression(m, d, 0)
    Amat = np.hstack([X, np.ones((m, 1))])
    nA = float(np.linalg.norm(Amat))
    row0, rows = AdaProx.sharding.shard_rows(m, world, rank)
    A = AdaProx.DeviceMatrix(Amat[row0:row0 + rows].copy(), dev=dev)
    if world > 1:
        AdaProx.sharding.attach_communicator(dev, dist)
        AdaProx.sharding.attach_p2p(dev, 4096, dist)
        A.set_shard(m, row0)
    Ac = AdaProx.Counting(A)
    h = AdaProx.Translate(AdaProx.NormL1(), -yv[row0:row0 + rows])
    kw = dict(f=AdaProx.Zero(), g=AdaProx.NormL1(10.0), h=h, A=Ac, eta=nA, t=1.0, tol=1e-5)
    AdaProx.adaptive_linesearch_primal_dual(np.zeros(d + 1), np.zeros(rows), maxit=20, **kw)
    Ac.mul_count = Ac.amul_count = 0
    dist.barrier()
    maxit = int(os.environ.get("PD_MAXIT", "2000"))
    x, y, it = AdaProx.adaptive_linesearch_primal_dual(np.zeros(d + 1), np.zeros(rows), maxit=maxit, **kw)
    info = AdaProx.last_solve_info()
    t = torch.tensor([info["solve_ms"]], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        passes = Ac.mul_count + Ac.amul_count
        print(json.dumps(dict(config=f"C3 LAD {m}x{d + 1} AdaPDM+ row-sharded over {world} GPU(s)", n_gpus=world, iterations=it,
                              device_ms=float(t[0]), us_per_iteration=1e3 * float(t[0]) / it, matrix_passes=passes,
                              aggregate_hbm_gbs=passes * m * (d + 1) * 8 / (float(t[0]) * 1e-3) / 1e9,
                              final_norm_res=info["final_norm_res"], collective=info["collective"])), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
