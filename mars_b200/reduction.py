"""Decomposition of groupby aggregation functions into pre / atomic-agg / post steps.

Mirror of what the reference's ReductionCompiler produces for the six functions on this path
(mars/dataframe/reduction/core.py:908-1250, probed in SURVEY.md Appendix A.3):

    sum   -> sum            | post: identity
    count -> count (merged with SUM of int64 counts, reduction/core.py:980-984)
    min   -> min,  max -> max
    mean  -> sum, count     | post: S / C                               (reduction/mean.py:26-33)
    var   -> count, sum(x**2), sum(x) | post: (Q - S**2 / C) / (C - ddof)  (reduction/var.py:38-48)
    std   -> as var, sqrt afterwards is not on this path (NotImplementedError)

The NamedTuples carry the reference's field names so that an executor written against the real operands
(``op.pre_funcs / op.agg_funcs / op.post_funcs``) works on both.  ``func`` members are small callable
objects with a symbolic ``kind`` -- the GPU executors dispatch on ``kind``; calling them on pandas objects
evaluates the same arithmetic in the reference's operation order (used for host-side checks).
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, NamedTuple, Optional, Sequence, Tuple


class ReductionPreStep(NamedTuple):
    input_key: str
    output_key: str
    columns: Optional[List[str]]
    func: Callable


class ReductionAggStep(NamedTuple):
    input_key: str
    raw_func_name: Optional[str]
    map_func_name: Optional[str]
    agg_func_name: Optional[str]
    custom_reduction: Optional[Any]
    output_key: str
    output_limit: int
    kwds: Dict[str, Any]


class ReductionPostStep(NamedTuple):
    input_keys: List[str]
    output_key: str
    func_name: str
    columns: Optional[List[str]]
    func: Callable


class ReductionSteps(NamedTuple):
    pre_funcs: List[ReductionPreStep]
    agg_funcs: List[ReductionAggStep]
    post_funcs: List[ReductionPostStep]


class PreFunc:
    """kind: 'identity' | 'square'"""

    def __init__(self, kind: str):
        self.kind = kind

    def __call__(self, x, gpu=None):
        return x if self.kind == "identity" else x ** 2

    def __repr__(self):
        return f"PreFunc({self.kind})"


class PostFunc:
    """kind: 'identity' | 'mean' (S, C) | 'var' (Q, S, C); evaluation order == the reference's generated
    source."""

    def __init__(self, kind: str, ddof: int = 1):
        self.kind = kind
        self.ddof = ddof

    def __call__(self, *inputs, gpu=None):
        if self.kind == "identity":
            return inputs[0]
        if self.kind == "mean":
            s, c = inputs
            return s / c
        if self.kind == "var":
            q, s, c = inputs
            if self.ddof == 0:
                return q / c - (s / c) ** 2
            t = s ** 2
            t = t / c
            u = q - t
            d = c - self.ddof
            return u / d
        raise NotImplementedError(self.kind)

    def __repr__(self):
        return f"PostFunc({self.kind})"


SUPPORTED_FUNCS = ("sum", "count", "min", "max", "mean", "var")

# func -> list of (pre kind, map func, agg func) atomic steps, in the reference's registration order
_DECOMP = {
    "sum": [("identity", "sum", "sum")],
    "count": [("identity", "count", "sum")],
    "min": [("identity", "min", "min")],
    "max": [("identity", "max", "max")],
    "mean": [("identity", "sum", "sum"), ("identity", "count", "sum")],
    "var": [("identity", "count", "sum"), ("square", "sum", "sum"), ("identity", "sum", "sum")],
}
_POST_KIND = {"sum": "identity", "count": "identity", "min": "identity", "max": "identity", "mean": "mean",
              "var": "var"}


def check_supported(func_name) -> str:
    if not isinstance(func_name, str) or func_name not in SUPPORTED_FUNCS:
        raise NotImplementedError(
            f"aggregation function {func_name!r} is outside the GPU groupby path "
            f"(supported: {', '.join(SUPPORTED_FUNCS)})")
    return func_name


def steps_of(func_name: str) -> List[Tuple[str, str, str]]:
    return list(_DECOMP[check_supported(func_name)])


_compiled: Dict[Any, "ReductionSteps"] = {}


def compile_reduction(funcs: Sequence[Tuple[Optional[str], str]], ddof: int = 1) -> ReductionSteps:
    """funcs: [(column or None, func name), ...] in user order -- (None, f) applies f to every value column
    (list form); (col, f) comes from the dict form.  Shared atomic steps are emitted once.  The result only
    depends on the arguments and is never modified by its users: compiled once per distinct call."""
    try:
        key = (tuple(funcs), ddof)
        hit = _compiled.get(key)
    except TypeError:
        key, hit = None, None
    if hit is None:
        hit = _compile_reduction(funcs, ddof)
        if key is not None:
            if len(_compiled) > 256:
                _compiled.clear()
            _compiled[key] = hit
    return hit


def _compile_reduction(funcs: Sequence[Tuple[Optional[str], str]], ddof: int = 1) -> ReductionSteps:
    pre: Dict[str, ReductionPreStep] = {}
    agg: Dict[Tuple[str, str], ReductionAggStep] = {}
    post: Dict[str, Dict[str, Any]] = {}
    referred: List[str] = []
    any_all_columns = False
    for col, f in funcs:
        check_supported(f)
        if col is None:
            any_all_columns = True
        elif col not in referred:
            referred.append(col)
    pre_cols = None if any_all_columns or not referred else list(referred)

    def pre_key(kind):
        key = "pre_" + kind
        if key not in pre:
            pre[key] = ReductionPreStep("pre_identity", key, pre_cols, PreFunc(kind))
        return key

    pre_key("identity")
    for col, f in funcs:
        keys = []
        for kind, map_f, agg_f in _DECOMP[f]:
            pk = pre_key(kind)
            ak = (pk, map_f)
            out_key = f"agg_{map_f}_{kind}"
            prev = agg.get(ak)
            # like the reference, raw_func_name is the last user function that registered the step
            agg[ak] = ReductionAggStep(pk, f, map_f, agg_f, None, out_key, 1,
                                       prev.kwds if prev else {"skipna": True, "axis": 0})
            keys.append(out_key)
        if f == "var":
            # reference input order: sum(x**2), sum(x), count  (Appendix A.3)
            keys = [keys[1], keys[2], keys[0]]
        entry = post.setdefault(f, dict(keys=keys, cols=[] if col is not None else None))
        if col is not None and entry["cols"] is not None:
            entry["cols"].append(col)
        elif col is None:
            entry["cols"] = None
    post_steps = []
    for f, entry in post.items():
        cols = entry["cols"]
        if cols is not None and pre_cols is not None and set(cols) == set(pre_cols):
            cols = None
        out_key = entry["keys"][0] if _POST_KIND[f] == "identity" else f"post_{f}"
        post_steps.append(ReductionPostStep(entry["keys"], out_key, f, cols, PostFunc(_POST_KIND[f], ddof)))
    return ReductionSteps(list(pre.values()), list(agg.values()), post_steps)
