"""Row-sharded AdaPGM over 2 GPUs (one process per GPU, NCCL all-reduce inside the
library) against the single-GPU solve and the CPU oracle.  Skipped on a 1-GPU box."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out, fused):
    sys.path.insert(0, ROOT)
    p2p = fused == "p2p"
    fused = "1" if p2p else fused
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank), ADAPROX_FUSED=fused)
    m2, n2, pf2 = (512, 4100, 40) if fused == "0" else (256, 20000, 200)     # fused: clusters of 3 CTAs
    import torch.distributed as dist
    import adaprox_b200 as AdaProx
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = AdaProx.Device(rank)
    AdaProx.set_default_device(dev)
    AdaProx.sharding.attach_communicator(dev, dist)
    if p2p:
        assert AdaProx.sharding.attach_p2p(dev, 20000, dist)   # in-kernel all-reduce over peer memory instead of NCCL
    m, n = 400, 1000
    P = AdaProx.synth.planted_lasso(m, n, 5, 0)
    Lf = AdaProx.synth.spectral_norm_sq(P["A"], iters=1000, tol=1e-15)
    row0, rows = AdaProx.sharding.shard_rows(m, world, rank)
    A = AdaProx.DeviceMatrix(P["A"][row0:row0 + rows], dev=dev)
    A.set_shard(m, row0)
    f = AdaProx.Counting(AdaProx.LinearLeastSquares(A, P["b"][row0:row0 + rows]))
    g = AdaProx.Counting(AdaProx.NormL1(1.0))
    log = []
    x, it = AdaProx.adaptive_proxgrad(np.zeros(n), f=f, g=g, rule=AdaProx.OurRule(gamma=1 / Lf), tol=1e-6, maxit=10000, log=log)
    # device-generated shard of a larger instance: every rank must reach the planted optimum
    Pd = AdaProx.generate_planted_lasso(m2, n2, pf2, 1, power_iters=60, row0=AdaProx.sharding.shard_rows(m2, world, rank)[0],
                                        rows=AdaProx.sharding.shard_rows(m2, world, rank)[1], dev=dev)
    log2 = []
    tol2, maxit2 = (1e-7, 30000) if fused == "0" else (0.0, 150)
    x2, it2 = AdaProx.adaptive_proxgrad(np.zeros(n2), f=AdaProx.LinearLeastSquares(Pd["A"], Pd["b"]), g=AdaProx.NormL1(1.0),
                                        rule=AdaProx.OurRule(gamma=1 / Pd["Lf"]), tol=tol2, maxit=maxit2, log=log2)
    info2 = AdaProx.last_solve_info()
    # the same instance unsharded on this GPU (same kernel family): the trajectories must agree to rounding
    Pf = AdaProx.generate_planted_lasso(m2, n2, pf2, 1, power_iters=60, dev=dev)
    log3 = []
    AdaProx.adaptive_proxgrad(np.zeros(n2), f=AdaProx.LinearLeastSquares(Pf["A"], Pf["b"]), g=AdaProx.NormL1(1.0),
                              rule=AdaProx.OurRule(gamma=1 / Pd["Lf"]), tol=0.0, maxit=100, log=log3)
    obj_sh = np.array([r["objective"] for r in log2[:100]]); obj_1 = np.array([r["objective"] for r in log3[:100]])
    gam_sh = np.array([r["gamma"] for r in log2[:100]]); gam_1 = np.array([r["gamma"] for r in log3[:100]])
    np.savez(out % rank, x=x, it=it, gam=np.array([r["gamma"] for r in log[:40]]), obj=log[-1]["objective"],
             counts=np.array([f.eval_count, f.grad_count, g.prox_count]), obj2=log2[-1]["objective"], opt2=Pd["optimum"],
             res2=log2[-1]["norm_res"], x2err=np.linalg.norm(x2 - Pd["x_star"]), launches=info2["kernel_launches"], passes=info2["matrix_passes"], it2=it2, collective=info2["collective"],
             obj_sh=obj_sh, obj_1=obj_1, gam_sh=gam_sh, gam_1=gam_1)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("fused", ["0", "1", "p2p"])
def test_sharded_adapgm_two_gpus(tmp_path, lasso_small, fused):
    """fused = "0": six split-phase launches per iteration (two sweeps over the shard) + ncclAllReduce; "1": the single-sweep
    cluster kernel in sweep-only mode + ncclAllReduce; "p2p": the same kernel with the all-reduce done inside it over NVLink
    peer memory (CUDA IPC mapped exchange buffers, system-scope flags)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from oracle import adaprox_oracle as O
    out = str(tmp_path / "rank%d.npz")
    mp.spawn(_worker, args=(2, 29400 + os.getpid() % 500 + 500 * ["0", "1", "p2p"].index(fused), out, fused), nprocs=2, join=True)
    R0, R1 = np.load(out % 0), np.load(out % 1)
    # replicated state stays in lock step: bit-identical iterates on both ranks
    assert np.array_equal(R0["x"], R1["x"]) and int(R0["it"]) == int(R1["it"])
    P = lasso_small
    logo = []
    xo, ito = O.adaptive_proxgrad(np.zeros(1000), f=O.LinearLeastSquares(P["A"], P["b"]), g=O.NormL1(1.0),
                                  rule=O.OurRule(gamma=1 / P["Lf"]), tol=1e-6, maxit=10000, log=logo)
    go = np.array([r["gamma"] for r in logo[:40]])
    assert np.max(np.abs(R0["gam"][:15] / go[:15] - 1)) < 1e-12
    assert np.max(np.abs(R0["gam"] / go - 1)) < 1e-9
    assert abs(float(R0["obj"]) - logo[-1]["objective"]) <= 1e-10 * abs(logo[-1]["objective"])
    assert abs(int(R0["it"]) - ito) <= max(2, 0.05 * ito)
    it = int(R0["it"])
    assert list(R0["counts"]) == [it + 1, it + 1, it]
    if fused == "0":
        assert float(R0["res2"]) <= 1e-7 and abs(float(R0["obj2"]) - float(R0["opt2"])) < 1e-9 * float(R0["opt2"])
        assert float(R0["x2err"]) < 1e-5
    assert np.allclose(R0["obj_sh"], R0["obj_1"], rtol=1e-9) and np.allclose(R0["gam_sh"][:20], R0["gam_1"][:20], rtol=1e-11)
    assert np.array_equal(R0["obj_sh"], R1["obj_sh"])
    assert int(R0["passes"]) == (2 if fused == "0" else 1)
    assert int(R0["collective"]) == (2 if fused == "p2p" else 1)
    if fused == "1":
        assert int(R0["launches"]) == 3 * (int(R0["it2"]) + 1)      # sweep kernel + E + F per gradient evaluation (ncclAllReduce between)
    if fused == "p2p":
        assert int(R0["launches"]) == 1                             # the whole sharded solve is one persistent launch per rank


# ---------------------------------------------------------------- row-sharded AdaPDM (LAD / sqrt-lasso)
def _worker_pd(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    import adaprox_b200 as AdaProx
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = AdaProx.Device(rank)
    AdaProx.set_default_device(dev)
    AdaProx.sharding.attach_communicator(dev, dist)
    assert AdaProx.sharding.attach_p2p(dev, 4096, dist)
    X, yv = AdaProx.synth.dense_regression(203, 10, 0)
    m = X.shape[0]
    Amat = np.hstack([X, np.ones((m, 1))])
    nA = float(np.linalg.norm(Amat))
    row0, rows = AdaProx.sharding.shard_rows(m, world, rank)
    res = {}
    for hname in ("l1", "l2"):
        A = AdaProx.DeviceMatrix(Amat[row0:row0 + rows].copy(), dev=dev)
        A.set_shard(m, row0)
        shift = -yv[row0:row0 + rows]
        h = AdaProx.Translate(AdaProx.NormL1() if hname == "l1" else AdaProx.NormL2(), shift)
        log = []
        x, y, it = AdaProx.adaptive_primal_dual(np.zeros(11), np.zeros(rows), f=AdaProx.Zero(), g=AdaProx.NormL1(0.1), h=h, A=AdaProx.Counting(A),
                                                rule=AdaProx.OurRule(t=1.0, norm_A=nA), tol=1e-6, maxit=300, log=log)
        info = AdaProx.last_solve_info()
        res[hname + "_x"] = x; res[hname + "_y"] = y; res[hname + "_it"] = it
        res[hname + "_gam"] = np.array([r["gamma"] for r in log]); res[hname + "_res"] = np.array([r["norm_res"] for r in log])
        res[hname + "_obj"] = np.array([r["objective"] for r in log]); res[hname + "_coll"] = info["collective"]
        res[hname + "_At"] = np.array([r["At_evals"] for r in log])
        # AdaPDM+ (linesearch on the estimate of |A|, src/AdaProx.jl:463-550) on the same shard
        A2 = AdaProx.DeviceMatrix(Amat[row0:row0 + rows].copy(), dev=dev)
        A2.set_shard(m, row0)
        log = []
        x, y, it = AdaProx.adaptive_linesearch_primal_dual(np.zeros(11), np.zeros(rows), f=AdaProx.Zero(), g=AdaProx.NormL1(0.1), h=AdaProx.Counting(h),
                                                           A=AdaProx.Counting(A2), eta=nA, t=1.0, tol=1e-6, maxit=300, log=log)
        res[hname + "_ls_x"] = x; res[hname + "_ls_it"] = it
        res[hname + "_ls_gam"] = np.array([r["gamma"] for r in log]); res[hname + "_ls_res"] = np.array([r["norm_res"] for r in log])
        res[hname + "_ls_At"] = np.array([r["At_evals"] for r in log]); res[hname + "_ls_ph"] = np.array([r["prox_h_evals"] for r in log])
    np.savez(out % rank, row0=row0, rows=rows, **res)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_adapdm_two_gpus(tmp_path):
    """AdaPDM (src/AdaProx.jl:312-364) with the linear map A row-sharded over 2 GPUs: the persistent kernel all-reduces A'y and the
    dual-side sums over NVLink peer memory itself.  Checked against the oracle on the whole problem."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import adaprox_b200 as AdaProx
    from oracle import adaprox_oracle as O
    out = str(tmp_path / "pd_rank%d.npz")
    mp.spawn(_worker_pd, args=(2, 31400 + os.getpid() % 500, out), nprocs=2, join=True)
    R0, R1 = np.load(out % 0), np.load(out % 1)
    X, yv = AdaProx.synth.dense_regression(203, 10, 0)
    m = X.shape[0]
    Amat = np.hstack([X, np.ones((m, 1))])
    nA = float(np.linalg.norm(Amat))
    for hname in ("l1", "l2"):
        assert int(R0[hname + "_coll"]) == 2
        assert np.array_equal(R0[hname + "_x"], R1[hname + "_x"]) and int(R0[hname + "_it"]) == int(R1[hname + "_it"])   # lock step
        ho = O.Translate(O.NormL1() if hname == "l1" else O.NormL2(), -yv)
        lo = []
        xo, yo, ito = O.adaptive_primal_dual(np.zeros(11), np.zeros(m), f=O.Zero(), g=O.NormL1(0.1), h=ho, A=O.Counting(Amat),
                                             rule=O.OurRule(t=1.0, norm_A=nA), tol=1e-6, maxit=300, log=lo)
        K = min(40, len(lo), len(R0[hname + "_gam"]))
        assert np.allclose(R0[hname + "_gam"][:K], [r["gamma"] for r in lo[:K]], rtol=1e-11)
        assert np.allclose(R0[hname + "_res"][:K], [r["norm_res"] for r in lo[:K]], rtol=1e-9)
        assert np.allclose(R0[hname + "_obj"][:K], [r["objective"] for r in lo[:K]], rtol=1e-10)
        assert list(R0[hname + "_At"][:K]) == [r["At_evals"] for r in lo[:K]]
        assert abs(int(R0[hname + "_it"]) - ito) <= max(3, 0.05 * ito)
        lo2 = []
        xo2, yo2, ito2 = O.adaptive_linesearch_primal_dual(np.zeros(11), np.zeros(m), f=O.Zero(), g=O.NormL1(0.1), h=O.Counting(ho), A=O.Counting(Amat),
                                                           eta=nA, t=1.0, tol=1e-6, maxit=300, log=lo2)
        assert np.array_equal(R0[hname + "_ls_x"], R1[hname + "_ls_x"])
        K2 = min(40, len(lo2), len(R0[hname + "_ls_gam"]))
        assert np.allclose(R0[hname + "_ls_gam"][:K2], [r["gamma"] for r in lo2[:K2]], rtol=1e-11)
        assert np.allclose(R0[hname + "_ls_res"][:K2], [r["norm_res"] for r in lo2[:K2]], rtol=1e-9)
        assert list(R0[hname + "_ls_At"][:K2]) == [r["At_evals"] for r in lo2[:K2]]          # same number of linesearch trials
        assert list(R0[hname + "_ls_ph"][:K2]) == [r["prox_h_evals"] for r in lo2[:K2]]
        assert abs(int(R0[hname + "_ls_it"]) - ito2) <= max(3, 0.05 * ito2)
        y = np.concatenate([R0[hname + "_y"], R1[hname + "_y"]])
        if int(R0[hname + "_it"]) == ito:
            assert np.allclose(R0[hname + "_x"], xo, rtol=1e-6, atol=1e-9) and np.allclose(y, yo, rtol=1e-6, atol=1e-9)


# ---------------------------------------------------------------- row-sharded dual SVM
def _svm_problem(factor=False):
    rng = np.random.default_rng(7)
    N, d = 90, 6
    X = rng.standard_normal((N, d))
    w = rng.standard_normal(d)
    y = np.sign(X @ w + 0.3 * rng.standard_normal(N))
    y[y == 0] = 1.0
    Z = y[:, None] * X
    if factor:
        return Z
    return Z @ Z.T, -np.ones(N), y                       # dual_svm/runme.jl:47-55: Q, q, labels


def _worker_svm(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    import adaprox_b200 as AdaProx
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = AdaProx.Device(rank)
    AdaProx.set_default_device(dev)
    AdaProx.sharding.attach_communicator(dev, dist)
    assert AdaProx.sharding.attach_p2p(dev, 4096, dist)
    Q, q, y = _svm_problem()
    N = Q.shape[0]
    row0, rows = AdaProx.sharding.shard_rows(N, world, rank)
    Qs = AdaProx.DeviceMatrix(Q[row0:row0 + rows].copy(), dev=dev)          # this rank's rows of the N x N matrix
    Qs.set_shard(N, row0)
    f = AdaProx.Counting(AdaProx.Quadratic(Qs, q))
    Amat = y[None, :].copy()
    nA = float(np.linalg.norm(Amat))
    log = []
    x, yy, it = AdaProx.adaptive_primal_dual(np.zeros(N), np.zeros(1), f=f, g=AdaProx.IndBox(0.0, 0.1), h=AdaProx.IndZero(),
                                             A=AdaProx.DeviceMatrix(Amat, dev=dev), rule=AdaProx.OurRule(t=1.0, norm_A=nA), tol=1e-6, maxit=3000, log=log)
    info = AdaProx.last_solve_info()
    # AdaPGM on the box-constrained quadratic alone (no equality constraint): the proximal-gradient entry point on a sharded Q
    log2 = []
    x2, it2 = AdaProx.adaptive_proxgrad(np.zeros(N), f=AdaProx.Quadratic(Qs, q), g=AdaProx.IndBox(0.0, 0.1), rule=AdaProx.OurRule(gamma=1e-2),
                                        tol=1e-7, maxit=2000, log=log2)
    # Gram form: Z = Dy X split by rows, Q never formed; u = Z'x is all-reduced inside the kernel (d doubles),
    # then the gradient rows and the value sums as above
    Z = _svm_problem(factor=True)
    Zs = AdaProx.DeviceMatrix(Z[row0:row0 + rows].copy(), dev=dev)
    Zs.set_shard(N, row0)
    f3 = AdaProx.Counting(AdaProx.QuadraticGram(Zs, q))
    log3 = []
    x3, y3, it3 = AdaProx.adaptive_primal_dual(np.zeros(N), np.zeros(1), f=f3, g=AdaProx.IndBox(0.0, 0.1), h=AdaProx.IndZero(),
                                               A=AdaProx.DeviceMatrix(Amat, dev=dev), rule=AdaProx.OurRule(t=1.0, norm_A=nA), tol=1e-6, maxit=3000, log=log3)
    coll3 = AdaProx.last_solve_info()["collective"]
    log4 = []
    x4, it4 = AdaProx.backtracking_proxgrad(np.zeros(N), f=AdaProx.QuadraticGram(Zs, q), g=AdaProx.IndBox(0.0, 0.1), gamma0=1.0, tol=1e-7, maxit=200, log=log4)
    np.savez(out % rank, x=x, y=yy, it=it, gam=np.array([r["gamma"] for r in log]), res=np.array([r["norm_res"] for r in log]),
             obj=np.array([r["objective"] for r in log]), coll=info["collective"], fe=f.eval_count, ge=f.grad_count,
             x2=x2, it2=it2, gam2=np.array([r["gamma"] for r in log2]), obj2=np.array([r["objective"] for r in log2]),
             x3=x3, it3=it3, gam3=np.array([r["gamma"] for r in log3]), obj3=np.array([r["objective"] for r in log3]), coll3=coll3,
             fe3=f3.eval_count, x4=x4, it4=it4, gam4=np.array([r["gamma"] for r in log4]), obj4=np.array([r["objective"] for r in log4]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_dual_svm_two_gpus(tmp_path):
    """dual_svm/runme.jl:47-59 with the N x N matrix Q row-sharded over 2 GPUs (x replicated): every rank computes its rows of
    Q x, the gradient rows and the two value sums are gathered inside the persistent kernel over NVLink peer memory."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from oracle import adaprox_oracle as O
    out = str(tmp_path / "svm_rank%d.npz")
    mp.spawn(_worker_svm, args=(2, 32400 + os.getpid() % 500, out), nprocs=2, join=True)
    R0, R1 = np.load(out % 0), np.load(out % 1)
    Q, q, y = _svm_problem()
    N = Q.shape[0]
    Amat = y[None, :].copy()
    nA = float(np.linalg.norm(Amat))
    assert int(R0["coll"]) == 2
    assert np.array_equal(R0["x"], R1["x"]) and int(R0["it"]) == int(R1["it"]) and np.array_equal(R0["x2"], R1["x2"])
    fo = O.Counting(O.Quadratic(Q, q))
    lo = []
    xo, yo, ito = O.adaptive_primal_dual(np.zeros(N), np.zeros(1), f=fo, g=O.IndBox(0.0, 0.1), h=O.IndZero(), A=Amat,
                                         rule=O.OurRule(t=1.0, norm_A=nA), tol=1e-6, maxit=3000, log=lo)
    K = min(40, len(lo), len(R0["gam"]))
    assert np.allclose(R0["gam"][:K], [r["gamma"] for r in lo[:K]], rtol=1e-11)
    assert np.allclose(R0["res"][:K], [r["norm_res"] for r in lo[:K]], rtol=1e-9)
    fin = np.isfinite([r["objective"] for r in lo[:K]])
    assert np.allclose(R0["obj"][:K][fin], np.array([r["objective"] for r in lo[:K]])[fin], rtol=1e-10)
    assert abs(int(R0["it"]) - ito) <= max(3, 0.05 * ito)
    assert (int(R0["fe"]), int(R0["ge"])) == (int(R0["it"]) + 1, int(R0["it"]) + 1)
    assert abs(fo.f(R0["x"]) - fo.f(xo)) <= 1e-7 * abs(fo.f(xo))
    lo2 = []
    xo2, ito2 = O.adaptive_proxgrad(np.zeros(N), f=O.Quadratic(Q, q), g=O.IndBox(0.0, 0.1), rule=O.OurRule(gamma=1e-2), tol=1e-7, maxit=2000, log=lo2)
    K2 = min(40, len(lo2), len(R0["gam2"]))
    assert np.allclose(R0["gam2"][:K2], [r["gamma"] for r in lo2[:K2]], rtol=1e-11)
    assert np.allclose(R0["obj2"][:K2], [r["objective"] for r in lo2[:K2]], rtol=1e-10)
    assert abs(int(R0["it2"]) - ito2) <= max(3, 0.05 * ito2)
    # Gram form on row shards of Z against the same dense-Q oracle runs (the two forms differ by rounding only)
    assert int(R0["coll3"]) == 2
    assert np.array_equal(R0["x3"], R1["x3"]) and int(R0["it3"]) == int(R1["it3"]) and np.array_equal(R0["x4"], R1["x4"])
    K3 = min(40, len(lo), len(R0["gam3"]))
    assert np.allclose(R0["gam3"][:K3], [r["gamma"] for r in lo[:K3]], rtol=1e-11)
    fin3 = np.isfinite([r["objective"] for r in lo[:K3]])
    assert np.allclose(R0["obj3"][:K3][fin3], np.array([r["objective"] for r in lo[:K3]])[fin3], rtol=1e-10)
    assert abs(int(R0["it3"]) - ito) <= max(3, 0.05 * ito) and int(R0["fe3"]) == int(R0["it3"]) + 1
    assert abs(fo.f(R0["x3"]) - fo.f(xo)) <= 1e-7 * abs(fo.f(xo))
    lo4 = []
    xo4, ito4 = O.backtracking_proxgrad(np.zeros(N), f=O.Quadratic(Q, q), g=O.IndBox(0.0, 0.1), gamma0=1.0, tol=1e-7, maxit=200, log=lo4)
    K4 = min(30, len(lo4), len(R0["gam4"]))
    assert np.allclose(R0["gam4"][:K4], [r["gamma"] for r in lo4[:K4]], rtol=1e-12)
    assert np.allclose(R0["obj4"][:K4], [r["objective"] for r in lo4[:K4]], rtol=1e-10)


# ---------------------------------------------------------------- row-sharded backtracking / Nesterov / aGRAAL baselines
def _worker_pg(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    import adaprox_b200 as AdaProx
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = AdaProx.Device(rank)
    AdaProx.set_default_device(dev)
    AdaProx.sharding.attach_communicator(dev, dist)
    assert AdaProx.sharding.attach_p2p(dev, 4096, dist)
    P = AdaProx.synth.planted_lasso(100, 300, 10, 0)
    Lf = AdaProx.synth.spectral_norm_sq(P["A"], iters=1000, tol=1e-15)
    g0, n, m = 1.0 / Lf, 300, 100
    row0, rows = AdaProx.sharding.shard_rows(m, world, rank)
    A = AdaProx.DeviceMatrix(P["A"][row0:row0 + rows].copy(), dev=dev)
    A.set_shard(m, row0)
    b = P["b"][row0:row0 + rows]
    res = {}

    def run(name, fn, **kw):
        f = AdaProx.Counting(AdaProx.LinearLeastSquares(A, b))
        log = []
        x, it = fn(np.zeros(n), f=f, g=AdaProx.NormL1(1.0), tol=1e-7, maxit=400, log=log, **kw)
        res[name + "_x"] = x; res[name + "_it"] = it; res[name + "_fe"] = f.eval_count; res[name + "_ge"] = f.grad_count
        res[name + "_gam"] = np.array([r["gamma"] for r in log]); res[name + "_obj"] = np.array([r["objective"] for r in log])
        res[name + "_res"] = np.array([r["norm_res"] for r in log]); res[name + "_coll"] = AdaProx.last_solve_info()["collective"]

    run("bt", AdaProx.backtracking_proxgrad, gamma0=g0, xi=1.5)
    run("btn", AdaProx.backtracking_nesterov, gamma0=g0)
    run("fn", AdaProx.fixed_nesterov, gamma=g0)
    run("ag", AdaProx.agraal, x0=np.random.default_rng(3).standard_normal(n), gamma0=g0)
    np.savez(out % rank, **res)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_pg_baselines_two_gpus(tmp_path):
    """backtracking_proxgrad / backtracking_nesterov / fixed_nesterov / agraal (src/AdaProx.jl:34-192) on a row-sharded lasso: the
    value sums of every trial and the gradients are combined inside the persistent kernel; same trial counts as the oracle."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import adaprox_b200 as AdaProx
    from oracle import adaprox_oracle as O
    out = str(tmp_path / "pg_rank%d.npz")
    mp.spawn(_worker_pg, args=(2, 33400 + os.getpid() % 500, out), nprocs=2, join=True)
    R0, R1 = np.load(out % 0), np.load(out % 1)
    P = AdaProx.synth.planted_lasso(100, 300, 10, 0)
    Lf = AdaProx.synth.spectral_norm_sq(P["A"], iters=1000, tol=1e-15)
    g0, n = 1.0 / Lf, 300
    cases = {"bt": (O.backtracking_proxgrad, dict(gamma0=g0, xi=1.5)), "btn": (O.backtracking_nesterov, dict(gamma0=g0)),
             "fn": (O.fixed_nesterov, dict(gamma=g0)), "ag": (O.agraal, dict(x0=np.random.default_rng(3).standard_normal(n), gamma0=g0))}
    for name, (fn, kw) in cases.items():
        assert int(R0[name + "_coll"]) == 2
        assert np.array_equal(R0[name + "_x"], R1[name + "_x"]) and int(R0[name + "_it"]) == int(R1[name + "_it"])
        fo = O.Counting(O.LinearLeastSquares(P["A"], P["b"]))
        lo = []
        xo, ito = fn(np.zeros(n), f=fo, g=O.NormL1(1.0), tol=1e-7, maxit=400, log=lo, **kw)
        K = min(30, len(lo), len(R0[name + "_gam"]))
        assert np.allclose(R0[name + "_gam"][:K], [r["gamma"] for r in lo[:K]], rtol=1e-11), name
        assert np.allclose(R0[name + "_obj"][:K], [r["objective"] for r in lo[:K]], rtol=1e-10), name
        assert np.allclose(R0[name + "_res"][:K], [r["norm_res"] for r in lo[:K]], rtol=1e-8), name
        assert abs(int(R0[name + "_it"]) - ito) <= max(3, 0.05 * ito), name
        if int(R0[name + "_it"]) == ito:
            assert (int(R0[name + "_fe"]), int(R0[name + "_ge"])) == (fo.eval_count, fo.grad_count), name    # same backtracking trials


# ---------------------------------------------------------------- row-sharded sparse logistic regression
def _worker_logreg(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    import time
    import scipy.sparse as sp
    import torch.distributed as dist
    import adaprox_b200 as AdaProx
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = AdaProx.Device(rank)
    AdaProx.set_default_device(dev)
    AdaProx.sharding.attach_communicator(dev, dist)
    res = {}
    for tag, (m, n, lo, hi, K) in dict(small=(600, 900, 10, 30, 60), c2=(20242, 47236, 40, 112, 120)).items():
        rp, ci, va, y = AdaProx.synth.sparse_logreg(m=m, n=n, seed=0, nnz_lo=lo, nnz_hi=hi)
        X = sp.csr_matrix((va, ci, rp), shape=(m, n))
        lam = 0.03 * AdaProx.synth.logreg_lambda_max(X, y)
        gam = 4 * m / (va @ va + m)
        row0, rows = AdaProx.sharding.shard_rows(m, world, rank)
        Xs = AdaProx.DeviceMatrix(X[row0:row0 + rows].tocsr(), dev=dev)
        Xs.set_shard(m, row0)
        f = AdaProx.Counting(AdaProx.LogisticLoss(Xs, y[row0:row0 + rows]))
        log = []
        t0 = time.perf_counter()
        x, it = AdaProx.adaptive_proxgrad(np.zeros(n + 1), f=f, g=AdaProx.NormL1(lam), rule=AdaProx.OurRule(gamma=gam), tol=0.0, maxit=K, log=log)
        info = AdaProx.last_solve_info()
        res[tag + "_x"] = x; res[tag + "_it"] = it
        res[tag + "_gam"] = np.array([r["gamma"] for r in log]); res[tag + "_obj"] = np.array([r["objective"] for r in log])
        res[tag + "_res"] = np.array([r["norm_res"] for r in log])
        res[tag + "_counts"] = np.array([f.eval_count, f.grad_count]); res[tag + "_ms"] = info["solve_ms"]; res[tag + "_coll"] = info["collective"]
    np.savez(out % rank, **res)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_sparse_logreg_two_gpus(tmp_path):
    """CSR logistic term split by rows over two GPUs (six split-phase launches + one ncclAllReduce of n + 2 doubles per iteration:
    X'r partials, loss sum, sum(p - y)); against the oracle on the whole problem, at a small shape and at configs[1]'s shape."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import scipy.sparse as sp
    import torch.multiprocessing as mp
    import adaprox_b200 as AdaProx
    from oracle import adaprox_oracle as O
    from oracle import drift
    out = str(tmp_path / "rank%d.npz")
    mp.spawn(_worker_logreg, args=(2, 27100 + os.getpid() % 500, out), nprocs=2, join=True)
    R0, R1 = np.load(out % 0), np.load(out % 1)
    for tag, (m, n, lo, hi, K) in dict(small=(600, 900, 10, 30, 60), c2=(20242, 47236, 40, 112, 120)).items():
        assert np.array_equal(R0[tag + "_x"], R1[tag + "_x"]) and np.array_equal(R0[tag + "_gam"], R1[tag + "_gam"])   # lock step
        rp, ci, va, y = AdaProx.synth.sparse_logreg(m=m, n=n, seed=0, nnz_lo=lo, nnz_hi=hi)
        X = sp.csr_matrix((va, ci, rp), shape=(m, n))
        lam = 0.03 * AdaProx.synth.logreg_lambda_max(X, y)
        gam = 4 * m / (va @ va + m)
        logo = []
        xo, ito = O.adaptive_proxgrad(np.zeros(n + 1), f=O.LogisticLoss(X, y), g=O.NormL1(lam), rule=O.OurRule(gamma=gam), tol=0.0, maxit=K, log=logo)
        go = np.array([r["gamma"] for r in logo])
        assert int(R0[tag + "_it"]) == ito == K
        # intrinsic drift of this instance: the oracle against itself with samples and features permuted (three samples)
        gps = []
        for sd in range(3):
            rng = np.random.default_rng(sd)
            pr, pc = rng.permutation(m), rng.permutation(n)
            lp = []
            O.adaptive_proxgrad(np.zeros(n + 1), f=O.LogisticLoss(X[pr][:, pc].tocsr(), y[pr]), g=O.NormL1(lam), rule=O.OurRule(gamma=gam), tol=0.0, maxit=40, log=lp)
            gps.append([r["gamma"] for r in lp])
        env = drift.perm_envelope(go[:40], gps)
        dd = np.abs(R0[tag + "_gam"][:40] / go[:40] - 1)
        assert np.all(dd <= np.maximum(1e-12, 20 * env)), (tag, float(dd.max()), float(env.max()))
        assert np.max(dd[:12]) < 1e-12, tag
        assert np.allclose(R0[tag + "_obj"][:25], [r["objective"] for r in logo[:25]], rtol=1e-10), tag
        assert np.allclose(R0[tag + "_res"][:25], [r["norm_res"] for r in logo[:25]], rtol=1e-8), tag
        assert np.linalg.norm(R0[tag + "_x"] - xo) <= 1e-6 * np.linalg.norm(xo), tag
        assert list(R0[tag + "_counts"]) == [K + 1, K + 1] and int(R0[tag + "_coll"]) == 1
    print("row-sharded CSR logreg, 2 GPUs, configs[1] shape: %.1f us per iteration" % (1e3 * float(R0["c2_ms"]) / 120))
