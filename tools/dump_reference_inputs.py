"""Writes the planted-lasso instance the device run uses (counter-based RNG, seed 0 unless given) as raw Float64 files for
julia/bench_reference.jl: A.f64 (column-major = Julia's layout), b.f64, x_star.f64 and meta.txt (m, n, lambda, gamma0 = 1/Lf,
optimum).  No GPU needed.   python tools/dump_reference_inputs.py --m 400 --n 1000 --out /tmp/c1"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adaprox_b200 as AdaProx  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=400)
    ap.add_argument("--n", type=int, default=1000)
    ap.add_argument("--pfactor", type=int, default=5)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    P = AdaProx.synth.planted_lasso(a.m, a.n, a.pfactor, a.seed)
    Lf = AdaProx.synth.spectral_norm_sq(P["A"], iters=1000, tol=1e-15)
    np.asfortranarray(P["A"]).T.tofile(os.path.join(a.out, "A.f64"))      # bytes of the column-major matrix
    P["b"].tofile(os.path.join(a.out, "b.f64"))
    P["x_star"].tofile(os.path.join(a.out, "x_star.f64"))
    with open(os.path.join(a.out, "meta.txt"), "w") as fh:
        fh.write(f"m={a.m}\nn={a.n}\nlambda=1.0\ngamma0={1.0 / Lf!r}\noptimum={P['optimum']!r}\nseed={a.seed}\n")
    print("wrote", a.out)


if __name__ == "__main__":
    main()
