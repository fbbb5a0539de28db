"""Cross-stencil fusion: a sequence of stencil calls -> ONE stencil IR.

The caller side of the hot path is a sequence of stencil calls per time step in which the output of
one call is the input of the next (reference caller: examples/cartesian/demo_burgers.ipynb cell 12).
Executed call by call every hand-over field makes a round trip through HBM.  `compose` turns the
sequence into a single stencil whose hand-over fields are *temporaries*: the streaming generator
(codegen_stream.py, `fuse_parallel_loops`) then keeps them in its register windows, so e.g. two
horizontal-diffusion steps move 12 B/cell instead of 24 (temporal blocking), and the fast-waves
chain drops every intermediate.

This is the cross-`StencilObject` extension of what the reference does inside one stencil with
`OnTheFlyMerging` (gtc/passes/oir_optimizations/horizontal_execution_merging.py) plus the extent
analysis of `compute_extents` (oir_optimizations/utils.py:250-330), restated here on the b200 IR:

* every call's IR is renamed into the program's name space (formal -> actual for fields and scalars,
  a per-call prefix for its temporaries and local scalars) and the vertical loops are concatenated;
* program fields listed in `intermediates` become temporaries;
* horizontal extents are recomputed backwards over the whole sequence: a producer is computed on
  the compute domain *extended by what its consumers read* (exactly gt4py's semantics inside one
  stencil), so `fused(A, B)` on domain D equals `A` on D grown by B's read extent followed by `B` on
  D — the input fields need the combined halo (reported in `field_info[...]["boundary"]`).  A
  hand-over field that is not listed in `intermediates` stays an argument and is additionally stored,
  on the grown domain, like a field written and re-read at an offset in a single gt4py stencil.

Restrictions (checked): intermediates are not read at K offsets / variable K, stencils with
horizontal regions are not composed (their extents are position dependent).
"""

from __future__ import annotations

import copy
from typing import Any, Dict, List, Optional, Sequence, Tuple

from . import ir as b2ir

Extent = List[List[int]]  # [[i0, i1], [j0, j1]], i0/j0 <= 0 <= i1/j1


def _union(a: Optional[Extent], b: Extent) -> Extent:
    if a is None:
        return [list(b[0]), list(b[1])]
    return [[min(a[0][0], b[0][0]), max(a[0][1], b[0][1])], [min(a[1][0], b[1][0]), max(a[1][1], b[1][1])]]


def _shift(e: Extent, di: int, dj: int) -> Extent:
    return [[e[0][0] + di, e[0][1] + di], [e[1][0] + dj, e[1][1] + dj]]


def _zero() -> Extent:
    return [[0, 0], [0, 0]]


def _with_zero(e: Extent) -> Extent:
    return [[min(e[0][0], 0), max(e[0][1], 0)], [min(e[1][0], 0), max(e[1][1], 0)]]


# ---- renaming ------------------------------------------------------------------------------------
def _rename_expr(node: Any, names: Dict[str, str]) -> None:
    def visit(e):
        if e["t"] in ("field", "scalar") and e["name"] in names:
            e["name"] = names[e["name"]]

    def stmt(s):
        t = s["t"]
        if t == "assign":
            b2ir.walk_exprs(s["left"], visit)
            b2ir.walk_exprs(s["right"], visit)
        elif t == "mask":
            b2ir.walk_exprs(s["mask"], visit)
            for b in s["body"]:
                stmt(b)
        elif t == "while":
            b2ir.walk_exprs(s["cond"], visit)
            for b in s["body"]:
                stmt(b)
        elif t == "hregion":
            raise NotImplementedError("b200 fuse: stencils with horizontal regions cannot be composed")
        else:
            raise ValueError(f"unknown stmt {t}")

    for s in node:
        stmt(s)


def _renamed_call(stencil: Dict[str, Any], binding: Dict[str, str], prefix: str) -> Dict[str, Any]:
    st = copy.deepcopy(stencil)
    names: Dict[str, str] = {}
    for p in st["params"]:
        names[p["name"]] = binding.get(p["name"], p["name"])
    for n in list(st["field_info"]) + list(st["parameter_info"]):
        names.setdefault(n, binding.get(n, n))
    for t in st["temporaries"]:
        names[t["name"]] = prefix + t["name"]
    for *_x, he in b2ir.iter_hes(st):
        for loc in he["locals"]:
            names[loc["name"]] = prefix + loc["name"]
    for p in st["params"]:
        p["name"] = names[p["name"]]
    for t in st["temporaries"]:
        t["name"] = names[t["name"]]
    for loop in st["loops"]:
        for c in loop.get("caches", []):
            c["name"] = names.get(c["name"], c["name"])
        for sec in loop["sections"]:
            for he in sec["hes"]:
                for loc in he["locals"]:
                    loc["name"] = names[loc["name"]]
                _rename_expr(he["body"], names)
    st["field_info"] = {names[n]: fi for n, fi in st["field_info"].items()}
    st["parameter_info"] = {names[n]: pi for n, pi in st["parameter_info"].items()}
    return st


# ---- extent analysis (restates compute_extents, oir_optimizations/utils.py:250-330, on the b200 IR) ----
def recompute_extents(stencil: Dict[str, Any]) -> Dict[str, Extent]:
    """Backward pass over the horizontal executions, the reference's rule: the block extent of a
    horizontal execution is the union of the extents at which the fields it writes are accessed later
    (API fields included: a field that is written and then read at an offset is computed — and stored —
    on the grown domain, as in a single gt4py stencil); every access then extends its field's extent by
    block extent + offset.  Sets he["extent"] and the temporaries' "extent"; returns field -> extent
    (for API fields: the halo the caller must provide)."""
    need: Dict[str, Extent] = {}
    # the reference's traversal (StencilExtentComputer, oir_optimizations/utils.py:276-300): vertical loops in reverse,
    # the sections of a loop in FORWARD order (no visitor reverses them), the horizontal executions of a section in reverse
    hes = [he for loop in reversed(stencil["loops"]) for sec in loop["sections"] for he in reversed(sec["hes"])]
    for he in hes:
        acc = b2ir.field_accesses(he["body"])
        ext = _zero()
        for a in acc:
            if a["write"] and a["name"] in need:
                ext = _union(ext, need[a["name"]])
        he["extent"] = ext
        for a in acc:
            di, dj = (0, 0) if a["write"] else b2ir.ij_offset(a["off"])
            need[a["name"]] = _union(need.get(a["name"]), _shift(ext, di, dj))
    for t in stencil["temporaries"]:
        t["extent"] = _with_zero(need.get(t["name"], _zero()))
    return {n: _with_zero(e) for n, e in need.items()}


def _access_kinds(stencil: Dict[str, Any]) -> Dict[str, str]:
    """READ / WRITE / READ_WRITE per API field, in program order (a field first written then read back is WRITE,
    like gt4py's AccessKind for outputs that are re-read: reference definitions.py:62-70)."""
    kinds: Dict[str, str] = {}
    for *_x, he in b2ir.iter_hes(stencil):
        for a in b2ir.field_accesses(he["body"]):
            k = kinds.get(a["name"])
            if a["write"]:
                kinds[a["name"]] = "READ_WRITE" if k in ("READ", "READ_WRITE") else "WRITE"
            elif k is None:
                kinds[a["name"]] = "READ"
    return kinds


def compose(name: str, calls: Sequence[Tuple[Dict[str, Any], Dict[str, str]]], *, intermediates: Sequence[str] = ()) -> Dict[str, Any]:
    """Fuse `calls` = [(stencil_ir, {formal argument: program name}), …] (executed in this order on the
    same compute domain) into one stencil IR.  Program fields in `intermediates` are produced and consumed
    inside the sequence only and become temporaries; all other fields / scalars are the arguments of
    the fused stencil (in order of first appearance)."""
    intermediates = list(intermediates)
    params: List[Dict[str, Any]] = []
    decl: Dict[str, Dict[str, Any]] = {}
    temporaries: List[Dict[str, Any]] = []
    loops: List[Dict[str, Any]] = []
    finfo_src: Dict[str, Dict[str, Any]] = {}
    pinfo: Dict[str, Any] = {}
    min_k = 0
    options: Dict[str, Any] = {}
    for n, (st, binding) in enumerate(calls):
        unknown = set(binding) - {p["name"] for p in st["params"]} - set(st["field_info"]) - set(st["parameter_info"])
        if unknown:
            raise ValueError(f"b200 fuse: call {n} ({st['name']}) has no argument(s) {sorted(unknown)}")
        r = _renamed_call(st, binding, f"c{n}_")
        for p in r["params"]:
            prev = decl.get(p["name"])
            if prev is None:
                decl[p["name"]] = p
                params.append(p)
            elif {k: prev[k] for k in prev if k != "name"} != {k: p[k] for k in p if k != "name"}:
                raise TypeError(f"b200 fuse: '{p['name']}' is used as {prev} and as {p}")
        temporaries += r["temporaries"]
        loops += r["loops"]
        for fname, fi in r["field_info"].items():
            if fi is None:
                finfo_src.setdefault(fname, None)
                continue
            if fname in intermediates and fi["access"] != "WRITE" and tuple(fi["boundary"][2]) != (0, 0):
                raise NotImplementedError(f"b200 fuse: intermediate '{fname}' is read at a K offset by call {n} ({st['name']})")
            prev = finfo_src.get(fname)
            if prev is None:
                finfo_src[fname] = copy.deepcopy(fi)
            else:  # keep the widest K boundary of all uses
                prev["boundary"][2] = [max(prev["boundary"][2][0], fi["boundary"][2][0]), max(prev["boundary"][2][1], fi["boundary"][2][1])]
        for pname, pi in r["parameter_info"].items():
            if pi is not None or pname not in pinfo:
                pinfo[pname] = pi
        min_k = max(min_k, int(st["domain_info"]["min_k"]))
        options.update(st.get("options", {}))
    for iname in intermediates:
        d = decl.get(iname)
        if d is None or d["t"] != "field":
            raise ValueError(f"b200 fuse: intermediate '{iname}' is not a field of the sequence")
        params.remove(d)
        temporaries.append({"name": iname, "dtype": d["dtype"], "dims": d["dims"], "data_dims": d["data_dims"], "extent": _zero()})
    fused = {
        "t": "stencil", "ir_version": b2ir.IR_VERSION, "name": name, "params": params, "temporaries": temporaries, "loops": loops,
        "variant": "fused", "options": options,
    }  # fmt: skip
    # an intermediate must be produced before it is consumed, everywhere it is consumed
    first: Dict[str, str] = {}
    for *_x, he in b2ir.iter_hes(fused):
        for a in b2ir.field_accesses(he["body"]):
            first.setdefault(a["name"], "write" if a["write"] else "read")
    for iname in intermediates:
        if first.get(iname) != "write":
            raise ValueError(f"b200 fuse: intermediate '{iname}' is read before the sequence writes it")
    need = recompute_extents(fused)
    kinds = _access_kinds(fused)
    field_info: Dict[str, Any] = {}
    for p in params:
        if p["t"] != "field":
            continue
        src = finfo_src.get(p["name"])
        if p["name"] not in kinds or src is None:
            field_info[p["name"]] = None if src is None else {**src, "access": "NONE"}
            continue
        e = need.get(p["name"], _zero())
        axes = src["axes"]
        field_info[p["name"]] = {
            "access": kinds[p["name"]],
            "boundary": [[-e[0][0] if "I" in axes else 0, e[0][1] if "I" in axes else 0],
                         [-e[1][0] if "J" in axes else 0, e[1][1] if "J" in axes else 0],
                         list(src["boundary"][2])],
            "axes": list(axes), "data_dims": list(src["data_dims"]), "dtype": src["dtype"],
        }  # fmt: skip
    fused["field_info"] = field_info
    fused["parameter_info"] = {p["name"]: pinfo.get(p["name"]) for p in params if p["t"] == "scalar"}
    fused["domain_info"] = {"min_k": min_k}
    return fused


def repeat(stencil: Dict[str, Any], times: int, *, carry: Tuple[str, str], name: Optional[str] = None) -> Dict[str, Any]:
    """Temporal blocking: `times` applications of one stencil in a single pass, the output `carry[1]` of an
    application being the input `carry[0]` of the next.  Arguments of the result: those of `stencil`."""
    src, dst = carry
    if times < 1:
        raise ValueError("times must be >= 1")
    calls, inter = [], []
    cur = src
    for n in range(times):
        out = dst if n == times - 1 else f"{dst}__step{n}"
        calls.append((stencil, {src: cur, dst: out}))
        if n < times - 1:
            inter.append(out)
        cur = out
    return compose(name or f"{stencil['name']}_x{times}", calls, intermediates=inter)


def ir_of(stencil: Any) -> Dict[str, Any]:
    """The b200 IR of a stencil given as an IR dict, a `B200Stencil`, or a gt4py `StencilObject` built with
    `backend="b200"` (its generated module records the IR file next to it in `.gt_cache`, backend.py)."""
    if isinstance(stencil, dict):
        return stencil
    ir = getattr(stencil, "ir", None)
    if isinstance(ir, dict):
        return ir
    run = getattr(type(stencil), "run", None)
    path = getattr(run, "__globals__", {}).get("_B200_IR")  # global of the generated module (backend.py)
    if path is None:
        raise TypeError(f"b200 fuse: {type(stencil).__name__} is not a b200 stencil")
    return b2ir.load_file(path)


def fuse_stencils(name: str, calls: Sequence[Tuple[Any, Dict[str, str]]], *, intermediates: Sequence[str] = (), options: Optional[Dict[str, Any]] = None):
    """`compose` for stencil objects: returns a callable `B200Stencil` that runs the whole sequence as one
    stencil (same call surface as a gt4py StencilObject: fields and scalars by program name,
    `origin=`, `domain=`)."""
    from .stencil import B200Stencil

    fused = compose(name, [(ir_of(s), dict(b)) for s, b in calls], intermediates=intermediates)
    return B200Stencil(fused, options, name=name)
