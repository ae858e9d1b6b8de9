"""2-rank probe of the in-kernel all-reduce: p2p vs NCCL trajectories and rank-to-rank agreement (torchrun)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ADAPROX_FUSED"] = "1"
import torch, torch.distributed as dist  # noqa: E402
import adaprox_b200 as AdaProx  # noqa: E402
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
dev = AdaProx.Device(local); AdaProx.set_default_device(dev)
AdaProx.sharding.attach_communicator(dev, dist)
AdaProx.sharding.attach_p2p(dev, 30000, dist)
m, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (400, 1000)
P = AdaProx.synth.planted_lasso(m, n, 5, 0)
Lf = AdaProx.synth.spectral_norm_sq(P["A"], iters=300)
row0, rows = AdaProx.sharding.shard_rows(m, world, rank)
A = AdaProx.DeviceMatrix(P["A"][row0:row0 + rows], dev=dev); A.set_shard(m, row0)
f = AdaProx.LinearLeastSquares(A, P["b"][row0:row0 + rows])
out = {}
for mode in ("p2p", "nccl", "p2p"):
    if mode == "nccl": os.environ["ADAPROX_NO_P2P"] = "1"
    else: os.environ.pop("ADAPROX_NO_P2P", None)
    log = []
    x, it = AdaProx.adaptive_proxgrad(np.zeros(n), f=f, g=AdaProx.NormL1(1.0), rule=AdaProx.OurRule(gamma=1 / Lf), tol=1e-6, maxit=300, log=log)
    g = np.array([r["gamma"] for r in log]); o = np.array([r["objective"] for r in log])
    info = AdaProx.last_solve_info()
    both = [None] * world
    dist.all_gather_object(both, (x, g, o, it))
    if rank == 0:
        dx = np.max(np.abs(both[0][0] - both[1][0])); k = min(len(both[0][1]), len(both[1][1]))
        dg = np.abs(both[0][1][:k] - both[1][1][:k]); first = int(np.argmax(dg > 0)) if np.any(dg > 0) else -1
        print(mode, "collective", info["collective"], "its", both[0][3], both[1][3], "rank diff x", dx, "first gamma mismatch at", first)
        if mode in out or True:
            out.setdefault(mode, []).append((g, o))
if rank == 0:
    gp, gn = out["p2p"][0][0], out["nccl"][0][0]
    k = min(len(gp), len(gn)); rel = np.abs(gp[:k] / gn[:k] - 1)
    print("p2p vs nccl gamma rel diff: first 10", rel[:10], "max", rel.max())
    print("p2p run 1 vs run 2 identical:", np.array_equal(out["p2p"][0][0], out["p2p"][1][0]))
dist.barrier(); dist.destroy_process_group()
