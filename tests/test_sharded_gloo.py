"""The N > 1 host logic on CPU: two gloo ranks each own a contiguous row block of
the planted lasso, run the oracle's AdaPGM loop with the A'r partials and the
value sums combined by ONE all-reduce of n + 2 doubles per gradient evaluation
(exactly what comm.inl does on the device), and must reproduce the unsharded
oracle run."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _ShardedLeastSquares:
    """LinearLeastSquares on a row shard; value and gradient are completed by a sum all-reduce."""

    def __init__(self, A_loc, b_loc, dist, torch):
        self.A, self.b, self.dist, self.torch = A_loc, b_loc, dist, torch
        self.n_allreduce = 0

    def eval_with_pullback(self, w):
        res = self.A @ w - self.b
        buf = np.empty(w.shape[0] + 2)
        buf[:-2] = self.A.T @ res
        buf[-2] = np.dot(res, res)
        buf[-1] = 0.0
        t = self.torch.from_numpy(buf)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        self.n_allreduce += 1
        nrm = np.sqrt(buf[-2])
        return 0.5 * nrm * nrm, (lambda: buf[:-2].copy())

    def __call__(self, w):
        return self.eval_with_pullback(w)[0]


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import adaprox_oracle as O
    import adaprox_b200

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = adaprox_b200.synth.planted_lasso(96, 200, 5, 2)
    Lf = adaprox_b200.synth.spectral_norm_sq(P["A"], iters=500)
    row0, rows = adaprox_b200.sharding.shard_rows(96, world, rank)
    # the shard can be generated in place from the counter-based RNG: same bits as the slice of the full matrix
    C_loc = adaprox_b200.synth.matrix_uniform_pm1(2, 96, 200, row0, rows)
    assert np.array_equal(C_loc * P["alpha"][None, :], P["A"][row0:row0 + rows])
    f = _ShardedLeastSquares(P["A"][row0:row0 + rows], P["b"][row0:row0 + rows], dist, torch)
    log = []
    x, it = O.adaptive_proxgrad(np.zeros(200), f=f, g=O.NormL1(1.0), rule=O.OurRule(gamma=1 / Lf), tol=1e-6, maxit=5000, log=log)
    # replicas must agree bit for bit (the all-reduce returns identical sums on every rank)
    t = torch.from_numpy(x.copy())
    gathered = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    if rank == 0:
        np.savez(out, x=x, it=it, gam=np.array([r["gamma"] for r in log[:40]]), obj=log[-1]["objective"], same=same,
                 n_allreduce=f.n_allreduce)
    dist.destroy_process_group()


def test_row_sharded_adapgm_world_size_2(tmp_path):
    import torch.multiprocessing as mp
    from oracle import adaprox_oracle as O
    import adaprox_b200

    out = str(tmp_path / "r0.npz")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    R = np.load(out)
    P = adaprox_b200.synth.planted_lasso(96, 200, 5, 2)
    Lf = adaprox_b200.synth.spectral_norm_sq(P["A"], iters=500)
    log = []
    x, it = O.adaptive_proxgrad(np.zeros(200), f=O.LinearLeastSquares(P["A"], P["b"]), g=O.NormL1(1.0),
                                rule=O.OurRule(gamma=1 / Lf), tol=1e-6, maxit=5000, log=log)
    assert bool(R["same"])
    assert np.allclose(R["gam"], [r["gamma"] for r in log[:40]], rtol=1e-11)
    assert abs(int(R["it"]) - it) <= max(2, 0.03 * it)
    assert abs(float(R["obj"]) - log[-1]["objective"]) < 1e-10 * abs(log[-1]["objective"])
    assert int(R["n_allreduce"]) == int(R["it"]) + 1                # one all-reduce per gradient evaluation, no other collective
    assert abs(float(R["obj"]) - P["optimum"]) < 1e-9 * P["optimum"]


# ---------------------------------------------------------------- attach_p2p: collective and failure-safe on every rank
class _FakeDev:
    """Stands in for adaprox_b200.Device: the export / attach calls of the peer exchange blocks, with scripted failures."""

    def __init__(self, rank, fail_export_on=None, fail_attach_on=None):
        self.rank, self.fail_export_on, self.fail_attach_on = rank, fail_export_on, fail_attach_on
        self.attached = None

    def p2p_export(self, n_max):
        if self.rank == self.fail_export_on:
            raise RuntimeError("cudaIpcGetMemHandle failed (scripted)")
        return bytes([self.rank]) * 64

    def p2p_attach(self, nranks, rank, handles):
        if self.rank == self.fail_attach_on:
            raise RuntimeError("cudaIpcOpenMemHandle failed (scripted)")
        assert len(handles) == 64 * nranks and handles[64 * rank] == rank
        self.attached = (nranks, rank)


def _worker_p2p(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    os.environ.pop("ADAPROX_NO_P2P", None)
    import adaprox_b200 as AdaProx
    dist.init_process_group("gloo", rank=rank, world_size=world)
    results = []
    for fe, fa in [(None, None), (1, None), (None, 0)]:
        os.environ.pop("ADAPROX_NO_P2P", None)
        dev = _FakeDev(rank, fe, fa)
        ok = AdaProx.sharding.attach_p2p(dev, 1000, dist)
        results.append((bool(ok), os.environ.get("ADAPROX_NO_P2P"), dev.attached))
    with open(out % rank, "w") as fh:
        fh.write(repr(results))
    dist.barrier()
    dist.destroy_process_group()


def test_attach_p2p_is_collective_and_falls_back(tmp_path):
    """A failure on ONE rank (export or attach) must not hang the others: every rank returns False and switches the
    library to ncclAllReduce (ADAPROX_NO_P2P=1); with no failure every rank attaches."""
    import torch.multiprocessing as mp
    out = str(tmp_path / "p2p_rank%d.txt")
    mp.spawn(_worker_p2p, args=(2, 30200 + os.getpid() % 500, out), nprocs=2, join=True)
    r0, r1 = eval(open(out % 0).read()), eval(open(out % 1).read())
    assert r0[0] == (True, None, (2, 0)) and r1[0] == (True, None, (2, 1))
    # export failed on rank 1: nobody attaches
    assert r0[1][:2] == (False, "1") and r1[1][:2] == (False, "1") and r0[1][2] is None and r1[1][2] is None
    # attach failed on rank 0: rank 1 did attach locally, but both report failure and fall back
    assert r0[2][:2] == (False, "1") and r1[2][:2] == (False, "1")
