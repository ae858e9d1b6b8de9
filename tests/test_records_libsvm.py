"""Host-side tooling of SURVEY section 8(f) rows 3-4: the record sink / find_best / is_logstep of
experiments/logging.jl and the LIBSVM reader of experiments/libsvm.jl.  No GPU."""
import json

import numpy as np
import pytest

import adaprox_b200 as AdaProx
from adaprox_b200 import records as R


def _julia_is_logstep(base, it):
    """logging.jl:13-17 evaluated in exact arithmetic (what the Julia code intends)."""
    scale = 0
    while base ** (scale + 1) <= it:
        scale += 1
    return it % (base ** scale) == 0


def test_is_logstep_matches_reference_definition():
    for base in (2, 10):
        for it in list(range(1, 1200)) + [5000, 10000, 12345, 20000, 99999, 100000]:
            assert R.is_logstep(base, it) == _julia_is_logstep(base, it)
    assert [it for it in range(1, 40) if R.is_logstep(10, it)] == [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 20, 30]
    assert R.is_logstep(10, 1000) and not R.is_logstep(10, 1100) and R.is_logstep(10, 2000)
    with pytest.raises(ValueError):
        R.is_logstep(10, 0)


def test_jsonl_sink_writes_reference_record_lines(tmp_path):
    path = str(tmp_path / "log.jsonl")
    rec1 = dict(method="AdaPGM", it=1, gamma=np.float64(0.5), sigma=np.float64(0.5), norm_res=np.float64(3.0),
                objective=np.float64(1.25), grad_f_evals=2, prox_g_evals=2, prox_h_evals=None, A_evals=None, At_evals=None, f_evals=2)
    rec2 = dict(rec1, it=2, objective=np.float64(np.inf), norm_res=np.float64(np.nan))
    with R.JsonlSink(path, mode="w") as sink:
        sink.append(rec1)
        sink.append(rec2)
        assert len(sink) == 2 and sink[1]["it"] == 2
    lines = open(path).read().splitlines()
    assert len(lines) == 2
    d1, d2 = json.loads(lines[0]), json.loads(lines[1])
    # key order = the order of the reference's @logmsg Record statement (src/AdaProx.jl:351)
    assert list(d1) == ["method", "it", "gamma", "sigma", "norm_res", "objective", "grad_f_evals", "prox_g_evals", "prox_h_evals",
                        "A_evals", "At_evals", "f_evals"]
    assert d1["gamma"] == 0.5 and d1["prox_h_evals"] is None and d1["f_evals"] == 2
    assert d2["objective"] is None and d2["norm_res"] is None          # JSON.jl writes non-finite numbers as null
    # keys = subset, reordered (logging.jl:24-26)
    path2 = str(tmp_path / "log2.jsonl")
    with R.JsonlSink(path2, keys=["it", "method", "objective"], mode="w") as sink:
        sink.append(rec1)
        with pytest.raises(KeyError):
            sink.append(dict(it=1))
    assert list(json.loads(open(path2).read().splitlines()[0])) == ["it", "method", "objective"]
    gb = R.read_jsonl(path)
    assert list(gb) == ["AdaPGM"] and [r["it"] for r in gb["AdaPGM"]] == [1, 2]


def _julia_find_best(gb, names, objective_key, objective_target, duration_key):
    """independent restatement of logging.jl:48-66 used as the checker"""
    names = list(names)
    best_name = names[0]
    best_duration = -1
    best_val = gb[best_name][-1][objective_key]
    if best_val <= objective_target:
        best_duration = max(r[duration_key] for r in gb[best_name])
    for name in names[1:]:
        duration = max(r[duration_key] for r in gb[name])
        val = gb[name][-1][objective_key]
        if val <= objective_target and (duration < best_duration or best_duration < 0):
            best_name, best_duration = name, duration
        elif best_duration < 0 and val < best_val:
            best_name, best_val = name, val
    return best_name


def test_find_best_follows_logging_jl():
    def run(final_obj, evals):
        return [dict(objective=10.0, grad_f_evals=1), dict(objective=final_obj, grad_f_evals=evals)]
    gb = {"a": run(1e-3, 50), "b": run(1e-7, 400), "c": run(1e-8, 300), "d": run(1e-2, 10)}
    assert R.find_best(gb, ["a", "b", "c", "d"], "objective", 1e-6, "grad_f_evals") == "c"      # reaches target, fewest evals
    assert R.find_best(gb, ["a", "d"], "objective", 1e-6, "grad_f_evals") == "a"               # nobody reaches it: smallest value
    assert R.find_best(gb, ["b"], "objective", 1e-6, "grad_f_evals") == "b"
    assert R.find_best(gb, ["c", "b"], "objective", 1e-6, lambda rows: [r["grad_f_evals"] for r in rows]) == "c"
    rng = np.random.default_rng(0)
    for _ in range(200):
        names = [f"m{k}" for k in range(int(rng.integers(1, 6)))]
        g2 = {nm: run(float(10.0 ** rng.integers(-9, 0)), int(rng.integers(1, 1000))) for nm in names}
        assert R.find_best(g2, names, "objective", 1e-6, "grad_f_evals") == _julia_find_best(g2, names, "objective", 1e-6, "grad_f_evals")
    with pytest.raises(ValueError):
        R.find_best(gb, [], "objective", 1e-6, "grad_f_evals")


def test_libsvm_reader(tmp_path):
    path = str(tmp_path / "toy.libsvm")
    with open(path, "w") as fh:
        fh.write("+1 1:0.5 3:2.0\n")
        fh.write("-1 2:1.5\n")
        fh.write("-1 1:1.0 1:0.25 4:-3\n")          # duplicate entry: summed by sparse(I, J, V)
        fh.write("+1\n")                            # a row without features
    X, y = R.load_libsvm_dataset(path)
    assert X.shape == (4, 4) and list(y) == [1.0, -1.0, -1.0, 1.0]
    dense = X.toarray()
    assert np.array_equal(dense, np.array([[0.5, 0, 2.0, 0], [0, 1.5, 0, 0], [1.25, 0, 0, -3.0], [0, 0, 0, 0]]))
    # label mapping (libsvm.jl:40-58): {-1, +1} -> (0, 1) as sparse_logreg/runme.jl uses
    X2, y2 = R.load_libsvm_dataset(path, labels=(0.0, 1.0))
    assert list(y2) == [1.0, 0.0, 0.0, 1.0]
    X3, y3 = R.load_libsvm_dataset(path, labels=(-1.0, 1.0))                  # already these values: untouched
    assert list(y3) == [1.0, -1.0, -1.0, 1.0]
    with pytest.raises(AssertionError):
        R.load_libsvm_dataset(path, labels=(1.0, 1.0))
    bad = str(tmp_path / "three.libsvm")
    open(bad, "w").write("0 1:1\n1 1:1\n2 1:1\n")
    with pytest.raises(AssertionError):
        R.load_libsvm_dataset(bad, labels=(0.0, 1.0))
    zero = str(tmp_path / "zero.libsvm")
    open(zero, "w").write("1 0:1\n")
    with pytest.raises(ValueError):
        R.load_libsvm_dataset(zero)


def test_package_exports():
    for nm in ("JsonlSink", "read_jsonl", "find_best", "is_logstep", "load_libsvm_dataset"):
        assert hasattr(AdaProx, nm)
