"""BASELINE configs[4] on N GPUs: the 256 lambdas of the batched lasso path are split over the ranks (A replicated, NO
collective on the data path -- SURVEY 8e); one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_path_multi.py

Prints one JSON line on rank 0: per-iteration time = max over ranks (device events), aggregate TFLOP/s over all ranks."""
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
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    dev = AdaProx.Device(local)
    AdaProx.set_default_device(dev)
    m, n, Lc = 16384, 8192, int(os.environ.get("C5_LAMBDAS", "256"))
    maxit = int(os.environ.get("C5_MAXIT", "100"))
    P = AdaProx.generate_planted_lasso(m, n, pfactor=5, seed=0, power_iters=30, dev=dev)       # the same instance on every rank
    f = AdaProx.LinearLeastSquares(P["A"], P["b"])
    lam_max = float(np.max(np.abs(P["A"].T @ P["b"].download())))
    lambdas = lam_max * (1e-3) ** (np.arange(Lc) / max(Lc - 1, 1))
    j0, j1 = (Lc * rank) // world, (Lc * (rank + 1)) // world
    mine = lambdas[j0:j1]
    AdaProx.adaptive_proxgrad_path(None, f=f, lambdas=mine, rule=AdaProx.OurRule(gamma=1 / P["Lf"]), tol=1e-6, maxit=5)   # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    X, its, info = AdaProx.adaptive_proxgrad_path(None, f=f, lambdas=mine, rule=AdaProx.OurRule(gamma=1 / P["Lf"]), tol=1e-6, maxit=maxit)
    evals = info["batched_evals"]
    t = torch.tensor([info["solve_ms"] / evals, float(evals)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_iter, ev = float(t[0]), float(t[1])
    if rank == 0:
        flop_iter = 4.0 * m * n * Lc
        print(json.dumps(dict(config=f"C5 lambda path {m}x{n} L={Lc} split over {world} GPU(s), {j1 - j0} lambdas per rank", n_gpus=world,
                              maxit=maxit, batched_evals=int(ev), ms_per_iteration=ms_iter, tflops_aggregate=flop_iter / (ms_iter * 1e-3) / 1e12,
                              lambda_iterations_per_s=Lc / (ms_iter * 1e-3), scaling="strong (256 lambdas fixed)", collective="none")), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
