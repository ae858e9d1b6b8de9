"""Drift curves for profiles/: the free-running AdaPGM stepsize sequence on configs[0] (planted lasso 400 x 1000, OurRule) computed by
  (x) the oracle in x87 extended precision (the yardstick), (a) the Float64 oracle, (b) the Float64 oracle with permuted columns (3 samples),
  (c) the device kernels through the C ABI: cluster-resident, single-sweep (forced), two-pass persistent grid kernel.
For every iteration k the relative distance of gamma_k from (x).  The device curves must stay inside the Float64 envelope (max of a, b) up to a
small factor: the loosened tolerances of the free-running parity tests are intrinsic to Float64, not to the CUDA path (oracle/drift.py).
    python tools/drift_curves.py > profiles/r02_drift_curves.json"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, ".")
import adaprox_b200 as AdaProx  # noqa: E402
from oracle import drift  # noqa: E402

K = 200
P = AdaProx.synth.planted_lasso(400, 1000, 5, 0)
Lf = AdaProx.synth.spectral_norm_sq(P["A"], iters=1000, tol=1e-15)
truth = drift.lasso_runs(P["A"], P["b"], 1.0, lambda O_: O_.OurRule(gamma=1 / Lf), K, nperm=3, seed=0)
ext = drift.series(truth["ext"])
out = {"instance": "planted lasso 400x1000 (configs[0]), AdaPGM OurRule gamma0 = 1/Lf, tol = 0, 200 iterations",
       "yardstick": "oracle/adaprox_oracle.py under precision(np.longdouble)",
       "iterations": list(range(1, K + 1)),
       "float64_oracle": drift.rel_to(ext, drift.series(truth["f64"])).tolist(),
       "float64_oracle_permuted_columns": [drift.rel_to(ext, drift.series(p)).tolist() for p in truth["perms"]],
       "float64_envelope_running_max": drift.envelope(truth).tolist()}
f, g = AdaProx.LinearLeastSquares(P["A"], P["b"]), AdaProx.NormL1(1.0)
for name, env in (("device_cluster_resident", {}), ("device_single_sweep", {"ADAPROX_FUSED": "1"}),
                  ("device_two_pass_grid", {"ADAPROX_RESIDENT": "0", "ADAPROX_GRIDRES": "0", "ADAPROX_FUSED": "0"})):
    os.environ.update(env)
    try:
        log = []
        AdaProx.adaptive_proxgrad(np.zeros(1000), f=f, g=g, rule=AdaProx.OurRule(gamma=1 / Lf), tol=0.0, maxit=K, log=log)
        out[name] = drift.rel_to(ext, [r["gamma"] for r in log]).tolist()
        out[name + "_matrix_passes"] = AdaProx.last_solve_info()["matrix_passes"]
    finally:
        for k in env:
            os.environ.pop(k, None)
env_ = np.array(out["float64_envelope_running_max"])
summary = {}
for name in ("device_cluster_resident", "device_single_sweep", "device_two_pass_grid"):
    d = np.array(out[name])
    summary[name] = {"max_ratio_to_envelope": float(np.max(d / np.maximum(env_, 1e-16))), "drift_at_10_50_100_200": [float(d[k - 1]) for k in (10, 50, 100, 200)]}
summary["envelope_at_10_50_100_200"] = [float(env_[k - 1]) for k in (10, 50, 100, 200)]
out["summary"] = summary
print(json.dumps(out))
