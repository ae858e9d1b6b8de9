"""Host-side logic that needs no GPU: the C-ABI library loads and exports every
symbol include/adaprox.h declares, struct layouts match, the row partition, the
counter-based RNG, the front-end's argument checking."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import adaprox_b200 as AdaProx
from adaprox_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "adaprox.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(adaprox_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = AdaProx.load()                                            # raises if the .so is missing
    declared = _declared_functions()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/adaprox.h but not exported"
    assert set(declared) == set(L.SYMBOLS), set(declared) ^ set(L.SYMBOLS)
    out = subprocess.run(["nm", "-D", "--defined-only", AdaProx.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (adaprox_[a-z0-9_]+)", out))
    assert set(declared) <= exported
    assert lib.adaprox_version() >= 100


def test_struct_layouts_match_the_header():
    # sizes computed from the header's field lists (natural alignment, x86-64)
    assert C.sizeof(L.Prox) == 4 + 4 + 8 * 3 + 8 * 3 == 56
    assert C.sizeof(L.Problem) == 4 + 4 + 8 + 8 + 8 + 56 + 56 + 8 + 8 + 8 == 168
    assert C.sizeof(L.Record) == 8 + 8 * 6 + 8 * 6 == 104
    assert C.sizeof(L.Result) == 8 + 4 + 4 + 8 * 6 + 8 + 8 * 3 + 8 + 8 + 8 + 8 == 128
    assert C.sizeof(L.Options) == 4 + 4 + 8 * 18 + 8 + 4 * 5 + 4 + 8 == 192


def test_enum_values_match_the_header():
    """every enumerator of include/adaprox.h has the same value in the ctypes mirror (and in the Julia shim's literal kinds)"""
    src = open(os.path.join(ROOT, "include", "adaprox.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    enums = {name: int(val) for name, val in re.findall(r"\bADAPROX_((?:F|P|S|RULE)_[A-Z0-9_]+)\s*=\s*(-?\d+)", src)}
    assert len(enums) == 8 + 5 + 8 + 4
    for name, val in enums.items():
        assert getattr(L, name) == val, name
    flags = {name: int(val) for name, val in re.findall(r"#define ADAPROX_(FLAG_[A-Z_]+)\s+(\d+)u", src)}
    assert flags and all(getattr(L, k) == v for k, v in flags.items())
    jl = open(os.path.join(ROOT, "julia", "AdaProxCUDA.jl")).read()
    for fn, kind in [("least_squares", "F_LEAST_SQUARES"), ("logistic", "F_LOGISTIC"), ("quadratic", "F_QUADRATIC"),
                     ("quadratic_gram", "F_QUADRATIC_GRAM"), ("cubic", "F_CUBIC"), ("worst_quadratic", "F_WORST_QUADRATIC")]:
        m = re.search(r"^%s\([^)]*\) = \(kind = (\d+)," % fn, jl, flags=re.M)
        assert m and int(m.group(1)) == enums[kind], fn


def test_julia_shim_struct_layouts_match_the_header():
    """julia/AdaProxCUDA.jl cannot be executed here (no Julia): at least its isbits structs must list the header's fields in the
    header's order with the matching Julia types, and every `ccall` must name an exported symbol."""
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "adaprox.h")).read(), flags=re.S)
    jl = open(os.path.join(ROOT, "julia", "AdaProxCUDA.jl")).read()
    ctype = {"int32_t": "Int32", "uint32_t": "UInt32", "int64_t": "Int64", "adaprox_id": "Id", "double": "Float64",
             "adaprox_prox": "CProx"}

    def c_fields(name):
        body = re.search(r"typedef struct \{([^{}]*)\} %s;" % name, hdr).group(1)
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            ty, names = decl.split(None, 1)
            out += [(n.strip(), ctype[ty]) for n in names.split(",")]
        return out

    def jl_fields(name):
        body = re.search(r"struct %s\n(.*?)\n(?:    \w+\(\) = new|end)" % name, jl, flags=re.S).group(1)
        return [(m.group(1), m.group(2)) for m in re.finditer(r"(\w+)::(\w+)", body)]

    for cname, jname in [("adaprox_prox", "CProx"), ("adaprox_problem", "CProblem"), ("adaprox_options", "COptions"),
                         ("adaprox_record", "CRecord"), ("adaprox_result", "CResult")]:
        cf, jf = c_fields(cname), jl_fields(jname)
        assert [t for _, t in cf] == [t for _, t in jf], (cname, cf, jf)
        assert [n for n, _ in cf] == [n for n, _ in jf], (cname, [n for n, _ in cf], [n for n, _ in jf])
    called = set(re.findall(r"ccall\(\(:(adaprox_\w+), lib\)", jl))
    assert called and called <= set(L.SYMBOLS), called - set(L.SYMBOLS)


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(AdaProx.AdaproxError):
        AdaProx.Device(0)


def test_stepsize_entry_point_runs_on_host():
    # adaprox_stepsize is pure host arithmetic (the same __host__ __device__ function the kernels call)
    from oracle import adaprox_oracle as O
    rng = np.random.default_rng(0)
    for rd, ro in [(AdaProx.OurRule(gamma=0.3), O.OurRule(gamma=0.3)),
                   (AdaProx.OurRule(t=2.0, norm_A=1.5, delta=0.02), O.OurRule(t=2.0, norm_A=1.5, delta=0.02)),
                   (AdaProx.MalitskyMishchenkoRule(0.3), O.MalitskyMishchenkoRule(0.3)),
                   (AdaProx.OurRulePlus(gamma=0.3), O.OurRulePlus(gamma=0.3))]:
        (_, _), st_d = AdaProx.stepsize(rd)
        (_, _), st_o = O.stepsize(ro)
        for _ in range(50):
            x1, x0 = rng.standard_normal(30), rng.standard_normal(30)
            g1, g0 = 2 * x1 + 0.1 * rng.standard_normal(30), 2 * x0 + 0.1 * rng.standard_normal(30)
            (go, so), st_o = O.stepsize(ro, st_o, x1, g1, x0, g0)
            dg, dx = g1 - g0, x1 - x0
            (gd, sd), st_d = AdaProx.stepsize(rd, st_d, np.dot(dg, dg), np.dot(dg, dx), np.dot(dx, dx))
            assert abs(gd - go) <= 1e-13 * abs(go) and abs(sd - so) <= 1e-13 * abs(so)


def test_shard_rows():
    for m, P in [(65536, 8), (65536, 1), (401, 3), (50000, 7), (8, 8), (17, 2)]:
        sh = AdaProx.sharding.all_shards(m, P)
        assert sh[0][0] == 0 and sum(r for _, r in sh) == m
        for (a, ra), (b, rb) in zip(sh[:-1], sh[1:]):
            assert a + ra == b and ra % 8 == 0
        assert max(r for _, r in sh) - min(r for _, r in sh) <= 8 + m % 8
    with pytest.raises(ValueError):
        AdaProx.sharding.shard_rows(4, 8, 0)


def test_counter_rng_is_random_access_and_shardable():
    s = AdaProx.synth
    full = s.matrix_uniform_pm1(3, 40, 30)
    part = s.matrix_uniform_pm1(3, 40, 30, row0=16, rows=8)
    assert np.array_equal(full[16:24], part)
    u = s.uniform01(0, 0, np.arange(200000))
    assert 0 <= u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 3e-3 and abs(u.var() - 1 / 12) < 2e-3
    assert s.bits64(0, 1, 5) != s.bits64(0, 2, 5) and s.bits64(0, 1, 5) != s.bits64(1, 1, 5)
    # known answers (shared with csrc/ops.cuh: stream_key / uniform01)
    with np.errstate(over="ignore"):
        assert int(s.bits64(0, 0, 0)) == int(s._mix(np.array([s.stream_key(0, 0)], dtype=np.uint64) + s._GOLD)[0])
    z = s.normal01(1, 1, np.arange(100000))
    assert abs(z.mean()) < 1e-2 and abs(z.std() - 1) < 1e-2


def test_front_end_argument_checks():
    with pytest.raises(ValueError):
        AdaProx.OurRule()
    with pytest.raises(AssertionError):
        AdaProx.fixed_nesterov(np.zeros(2), f=None, g=None)        # src/AdaProx.jl:104
    with pytest.raises(AssertionError):
        AdaProx.adaptive_linesearch_primal_dual(np.zeros(2), np.zeros(2), f=None, g=None, h=None, A=None, Theta=1.0, delta=0.5)
    r = AdaProx.OurRule(t=2.0, norm_A=4.0)
    assert r.gamma == 1 / (2 * 1.2 * 2.0 * 4.0)                     # :244


def test_generators_shapes():
    rp, ci, va, y = AdaProx.synth.sparse_logreg(m=50, n=70, seed=1, nnz_lo=3, nnz_hi=9)
    assert rp[0] == 0 and rp[-1] == len(ci) == len(va) and set(np.unique(y)) <= {0.0, 1.0}
    for i in range(50):
        row = va[rp[i]:rp[i + 1]]
        assert abs(np.dot(row, row) - 1) < 1e-12 and np.all(np.diff(ci[rp[i]:rp[i + 1]]) > 0)
    X, yy = AdaProx.synth.dense_classification(40, 6, 0)
    assert X.shape == (40, 6) and set(np.unique(yy)) <= {-1.0, 1.0}
