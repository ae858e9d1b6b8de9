"""Record sink, log-step filter and best-run selection compatible with the reference's experiment tooling
(experiments/logging.jl:13-66), and the LIBSVM text reader (experiments/libsvm.jl:3-61).

The solvers of this package take ``log=<list-like>`` and append one dict per iteration with the keys of the
reference's ``@logmsg Record`` statements (src/AdaProx.jl:56,135,351,539,621), in the same order.  ``JsonlSink`` is
such a list-like that writes each record as one JSON line, which is exactly what ``get_logger`` produces
(logging.jl:19-27: ``println(io, JSON.json(args.kwargs))``), so the reference's ``plot_*`` readers
(``eachline(path) .|> JSON.parse``, lasso/runme.jl:164) consume files written from device runs unchanged.
Host-side only: no GPU and no oracle involved.
"""
from __future__ import annotations

import json
import math
from typing import Callable, Iterable, Mapping, Sequence

import numpy as np


def is_logstep(base: int, it: int) -> bool:
    """experiments/logging.jl:13-17, restated as written: ``scale = floor(Int, log(base, it)); step = base^scale;
    mod(it, step) == 0`` with Julia's ``log(b, x) = log(x) / log(b)`` in Float64.  (The floating-point logarithm of an
    exact power can land just below the integer -- log(10, 1000) = 2.9999999999999996 gives step 100 instead of 1000 --
    but an exact power is a multiple of the smaller step too, so the answer is the same as with the exact integer power.)"""
    if it < 1:
        raise ValueError("is_logstep: it must be >= 1")               # Julia: DomainError / InexactError from floor(Int, -Inf)
    scale = math.floor(math.log(it) / math.log(base))
    step = base ** scale
    return it % step == 0


def _jsonable(v):
    """JSON.jl semantics for the values that occur in records: non-finite floats and ``nothing`` become null."""
    if v is None:
        return None
    if isinstance(v, (np.floating, float)):
        v = float(v)
        return v if math.isfinite(v) else None
    if isinstance(v, (np.integer, int)) and not isinstance(v, bool):
        return int(v)
    return v


class JsonlSink:
    """``log=JsonlSink(path, keys=None)``: one JSON object per record (logging.jl:19-27).  ``keys`` selects and orders a
    subset of the fields like ``args.kwargs[keys]``.  Also keeps the records in memory (``.records``) unless
    ``keep=False``.  ``console_base`` prints the records that pass ``is_logstep`` (logging.jl:33-39)."""

    def __init__(self, path: str, keys: Sequence[str] | None = None, mode: str = "a", keep: bool = True,
                 console_base: int | None = None):
        self.path, self.keys, self.keep, self.console_base = path, (list(keys) if keys is not None else None), keep, console_base
        self.records: list[dict] = []
        self._fh = open(path, mode)

    def append(self, rec: Mapping) -> None:
        if self.keys is not None:
            missing = [k for k in self.keys if k not in rec]
            if missing:
                raise KeyError(f"record has no field(s) {missing}")            # Julia: kwargs[keys] throws as well
            out = {k: _jsonable(rec[k]) for k in self.keys}
        else:
            out = {k: _jsonable(v) for k, v in rec.items()}
        self._fh.write(json.dumps(out) + "\n")
        if self.keep:
            self.records.append(dict(rec))
        if self.console_base is not None and is_logstep(self.console_base, int(rec["it"])):
            print(" ".join(f"{k}={v}" for k, v in out.items()), flush=True)

    def extend(self, recs: Iterable[Mapping]) -> None:
        for r in recs:
            self.append(r)

    def __len__(self) -> int:
        return len(self.records)

    def __getitem__(self, i):
        return self.records[i]

    def __iter__(self):
        return iter(self.records)

    def flush(self) -> None:
        self._fh.flush()

    def close(self) -> None:
        if not self._fh.closed:
            self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def read_jsonl(path: str) -> dict[str, list[dict]]:
    """Records of a JSONL file grouped by ``method`` in first-appearance order (the ``groupby(df, :method)`` of the
    reference's plot scripts)."""
    groups: dict[str, list[dict]] = {}
    with open(path) as fh:
        for line in fh:
            line = line.strip()
            if not line:
                continue
            rec = json.loads(line)
            groups.setdefault(rec.get("method"), []).append(rec)
    return groups


def _duration(rows: Sequence[Mapping], key) -> float:
    """logging.jl:44-46: ``maximum(df[!, key])`` or ``maximum(fun(df))``."""
    if callable(key):
        vals = key(rows)
    else:
        vals = [r[key] for r in rows]
    return max(vals)


def find_best(gb: Mapping[str, Sequence[Mapping]], names: Iterable[str], objective_key: str, objective_target: float,
              duration_key: str | Callable) -> str:
    """logging.jl:48-66, statement for statement: among ``names`` pick the run that reaches ``objective_target``
    (last value of ``objective_key`` <= target) in the smallest duration; if none does, the one with the smallest
    last value.  ``None`` values (JSON null: a non-finite objective) compare as +Inf."""
    def last(name):
        v = gb[name][-1][objective_key]
        return math.inf if v is None else v
    it = iter(names)
    try:
        best_name = next(it)
    except StopIteration:
        raise ValueError("find_best: no names") from None
    best_duration = -1
    best_val = last(best_name)
    if best_val <= objective_target:
        best_duration = _duration(gb[best_name], duration_key)
    for name in it:
        duration = _duration(gb[name], duration_key)
        val = last(name)
        if val <= objective_target and (duration < best_duration or best_duration < 0):
            best_name = name
            best_duration = duration
        elif best_duration < 0 and val < best_val:
            best_name = name
            best_val = val
    return best_name


def load_libsvm_dataset(file_path: str, dtype=np.float64, labels: Sequence[float] | None = None):
    """experiments/libsvm.jl:3-61.  Returns ``(X, y)`` with ``X`` a scipy CSR matrix (the reference builds a
    ``SparseMatrixCSC`` with ``sparse(rows, cols, vals)``: 1-based column indices, duplicate entries summed, as many
    columns as the largest index seen) and ``y`` the first token of every line.  With ``labels=(l0, l1)`` the two
    label values found in the file are mapped to l0 (smaller) and l1 (larger) unless they already are these values
    (libsvm.jl:40-58).  Upload with ``DeviceMatrix(X)`` / ``LogisticLoss(X, y)``."""
    import scipy.sparse as sp
    if labels is not None:
        if len(labels) != 2:
            raise AssertionError("length(labels) == 2")
        if labels[0] == labels[1]:
            raise AssertionError("labels[1] != labels[2]")
    y, rows, cols, vals = [], [], [], []
    with open(file_path) as fh:
        for row, line in enumerate(fh):
            tokens = line.strip().split(" ")
            if tokens == [""]:
                raise ValueError(f"{file_path}:{row + 1}: empty line")         # Julia: parse(T, "") throws
            y.append(float(tokens[0]))
            for token in tokens[1:]:
                col_str, val_str = token.split(":")
                col = int(col_str)
                if col < 1:
                    raise ValueError(f"{file_path}:{row + 1}: column index {col} (LIBSVM indices are 1-based)")
                rows.append(row)
                cols.append(col - 1)
                vals.append(float(val_str))
    y = np.asarray(y, dtype=dtype)
    m = len(y)
    n = (max(cols) + 1) if cols else 0
    X = sp.coo_matrix((np.asarray(vals, dtype=dtype), (np.asarray(rows, dtype=np.int64), np.asarray(cols, dtype=np.int64))),
                      shape=(m, n)).tocsr()                                    # duplicates are summed, like sparse(I, J, V)
    X.sort_indices()
    if labels is not None:
        uniq = np.unique(y)
        if len(uniq) != 2:
            raise AssertionError("length(unique(y)) == 2")
        y0, y1 = uniq[0], uniq[1]
        l0, l1 = labels
        if not (y0 in labels and y1 in labels):
            was0, was1 = (y == y0), (y == y1)
            y[was0] = l0
            y[was1] = l1
        if not np.all(np.isin(y, np.asarray(labels, dtype=dtype))):
            raise AssertionError("all(in(v, labels) for v in y)")
    return X, y
