"""CPU oracle for the b200 backend — TEST INFRASTRUCTURE ONLY, never on the product path.

A NumPy *interpreter* of the b200 stencil IR (`gt4py_b200/ir.py`) that restates, node for node, the
execution model of the reference's `numpy` backend:

* whole-array, statement-major execution of every assignment over the horizontal block
  `[i:I, j:J]` given by the block extent         (reference: gtc/numpy/npir_codegen.py:298-318)
* PARALLEL  : each statement over the full `[k:K]` range of its section
  FORWARD   : `for k_ in range(k, K)`, BACKWARD: `for k_ in range(K-1, k-1, -1)`,
  all horizontal blocks of the section inside one level        (npir_codegen.py:243-248, 269-296)
* interval bounds `START+o -> o`, `END+o -> nK+o`                   (npir_codegen.py:227-241)
* masks  -> `lhs = where(mask, rhs, lhs)`, nested masks AND-ed       (gtc/numpy/oir_to_npir.py:148-174)
* while  -> `while np.any(cond): body` with the body masked by cond  (oir_to_npir.py:176-185,
                                                                       npir_codegen.py:252-267)
* horizontal regions -> statement restricted to (block extent ∩ region)
                                                  (oir_to_npir.py:187-198, passes/horizontal_masks.py:90-112)
* origin shift of every access, broadcast of missing axes, variable-K reads clipped to the
  field's K range                                            (cartesian/utils/field.py:19-58)
* temporaries of shape `domain + extent padding`, origin `-extent.lower`   (oir_to_npir.py:42-56)
* arithmetic through the same NumPy/SciPy ufuncs              (gtc/ufuncs.py:15-93, gtc/common.py:910-993)
* `np.errstate(divide/over/under/invalid = 'ignore')`               (npir_codegen.py:355-359)

Parity pin: `tests/golden/*.npz` hold outputs of the *reference itself* (its `numpy` backend
imported from /root/reference in the build container by `tools/make_golden.py`); the CPU test-suite
checks this interpreter against every one of them bit for bit.

Only `tests/`, `__graft_entry__.smoke()`, the cpu-baseline / `--impl reference` legs of `bench.py`
(and of its sibling for the other configs, `tools/bench_workloads.py`) and the golden-vector generator
`tools/make_golden.py` may import this module; nothing under `gt4py_b200/` does.
"""

from __future__ import annotations

from typing import Any, Dict, Optional, Tuple

import numpy as np

try:  # the reference uses scipy.special when available (gtc/ufuncs.py:15-29)
    from scipy.special import erf as _erf, erfc as _erfc, gamma as _gamma
except ImportError:  # pragma: no cover
    import math

    _gamma = np.vectorize(math.gamma)
    _erf = np.vectorize(math.erf)
    _erfc = np.vectorize(math.erfc)


def _round_away_from_zero(x):
    return np.copysign(np.floor(np.abs(x) + 0.5), x)


_NP_DTYPE = {
    "bool": np.bool_,
    "int8": np.int8,
    "int16": np.int16,
    "int32": np.int32,
    "int64": np.int64,
    "float32": np.float32,
    "float64": np.float64,
}

_NATIVE = {
    "abs": np.abs,
    "min": np.minimum,
    "max": np.maximum,
    "mod": np.remainder,
    "sin": np.sin,
    "cos": np.cos,
    "tan": np.tan,
    "arcsin": np.arcsin,
    "arccos": np.arccos,
    "arctan": np.arctan,
    "sinh": np.sinh,
    "cosh": np.cosh,
    "tanh": np.tanh,
    "arcsinh": np.arcsinh,
    "arccosh": np.arccosh,
    "arctanh": np.arctanh,
    "sqrt": np.sqrt,
    "pow": np.power,
    "exp": np.exp,
    "log": np.log,
    "log10": np.log10,
    "gamma": _gamma,
    "cbrt": np.cbrt,
    "isfinite": np.isfinite,
    "isinf": np.isinf,
    "isnan": np.isnan,
    "floor": np.floor,
    "ceil": np.ceil,
    "trunc": np.trunc,
    "erf": _erf,
    "erfc": _erfc,
    "round": np.round,
    "round_away_from_zero": _round_away_from_zero,
    "int32": np.int32,
    "int64": np.int64,
    "float32": np.float32,
    "float64": np.float64,
}


class _Field:
    """An array plus its origin, addressed in domain coordinates (cartesian/utils/field.py:15-74)."""

    def __init__(self, array: np.ndarray, origin: Tuple[int, ...], dims):
        it = iter(range(3))
        self.idx_to_data = [next(it) if d else None for d in dims]
        n_api = sum(bool(d) for d in dims)
        shape = [array.shape[i] if i is not None else 1 for i in self.idx_to_data] + list(array.shape[n_api:])
        self.view = array.reshape(shape) if array.flags.c_contiguous else _reshape_view(array, self.idx_to_data)
        self.origin = tuple(origin)

    def key(self, i_sl, j_sl, k_idx, data_index):
        if isinstance(k_idx, slice):
            key = []
            for axis, sl in enumerate((i_sl, j_sl, k_idx)):
                src = self.idx_to_data[axis]
                if src is None:
                    key.append(slice(None))
                else:
                    o = self.origin[src]
                    key.append(slice(sl.start + o, sl.stop + o))
            return tuple(key) + tuple(data_index)

        # variable K offset: integer index arrays, K clipped to the array (utils/field.py:54-58)
        def axis_index(axis, sl, shape):
            src = self.idx_to_data[axis]
            if src is None:
                return np.zeros(shape, dtype=np.int64)
            o = self.origin[src]
            return np.arange(sl.start + o, sl.stop + o).reshape(shape)

        ii = axis_index(0, i_sl, (-1, 1, 1))
        jj = axis_index(1, j_sl, (1, -1, 1))
        srck = self.idx_to_data[2]
        if srck is None:
            kk = np.zeros((1, 1, 1), dtype=np.int64)
        else:
            kk = np.clip(np.asarray(k_idx) + self.origin[srck], 0, self.view.shape[2] - 1)
        return (ii, jj, kk) + tuple(data_index)

    def get(self, i_sl, j_sl, k_idx, data_index=()):
        return self.view[self.key(i_sl, j_sl, k_idx, data_index)]

    def set(self, i_sl, j_sl, k_idx, data_index, value):
        self.view[self.key(i_sl, j_sl, k_idx, data_index)] = value


def _reshape_view(array, idx_to_data):
    # insert size-1 axes for missing dimensions without copying (works for any strides)
    out = array
    for axis, src in enumerate(idx_to_data):
        if src is None:
            out = np.expand_dims(out, axis)
    return out


class _Ctx:
    def __init__(self, fields, params, domain):
        self.fields: Dict[str, _Field] = fields
        self.params = params
        self.nI, self.nJ, self.nK = domain
        self.locals: Dict[str, np.ndarray] = {}


def _literal(node):
    dt = _NP_DTYPE[node["dtype"]]
    v = node["value"]
    if node["dtype"] == "bool":
        return np.bool_(v in ("True", "true", "1"))
    if node["dtype"].startswith("int"):
        return dt(int(v))
    return dt(float(v))


def _eval(node, ctx: _Ctx, reg) -> Any:
    """Evaluate an expression over region reg=(i0,i1,j0,j1,k0,k1) -> array (I,J,K) or scalar."""
    t = node["t"]
    i0, i1, j0, j1, k0, k1 = reg
    if t == "field":
        name = node["name"]
        off = node["off"]
        data_index = tuple(int(_eval(d, ctx, reg)) for d in node.get("data_index", []))
        if name in ctx.locals:
            arr = ctx.locals[name]
            return arr[_local_key(ctx, reg)]
        f = ctx.fields[name]
        if isinstance(off, dict):
            if "vk" in off:
                dk = _eval(off["vk"], ctx, reg)
                lk = np.arange(k0, k1)[None, None, :]
                kidx = (lk + dk).astype(np.int64)
                kidx = np.broadcast_to(kidx, np.broadcast_shapes(kidx.shape, (1, 1, k1 - k0)))
                return f.get(slice(i0, i1), slice(j0, j1), kidx, data_index)
            raise NotImplementedError(
                "Absolute K indexation (e.g. `field.at(...)`) is an experimental feature and not "
                "yet implemented for the `numpy` backend."
            )
        di, dj, dk = off
        return f.get(slice(i0 + di, i1 + di), slice(j0 + dj, j1 + dj), slice(k0 + dk, k1 + dk), data_index)
    if t == "scalar":
        name = node["name"]
        if name in ctx.locals:
            return ctx.locals[name][_local_key(ctx, reg)]
        return ctx.params[name]
    if t == "lit":
        return _literal(node)
    if t == "iter":
        if node["axis"] != "K":
            raise ValueError(f"Axis {node['axis']} cannot be accessed, only K.")
        return np.arange(k0, k1, dtype=int)[None, None, :] + np.zeros((i1 - i0, j1 - j0, 1), dtype=int)
    if t == "cast":
        v = _eval(node["expr"], ctx, reg)
        dt = _NP_DTYPE[node["dtype"]]
        return v.astype(dt) if isinstance(v, np.ndarray) else dt(v)
    if t == "unary":
        v = _eval(node["expr"], ctx, reg)
        op = node["op"]
        if op == "not":
            return np.bitwise_not(v)
        return -v if op == "-" else +v
    if t == "binary":
        a = _eval(node["left"], ctx, reg)
        b = _eval(node["right"], ctx, reg)
        op = node["op"]
        if op == "+":
            return a + b
        if op == "-":
            return a - b
        if op == "*":
            return a * b
        if op == "/":
            return a / b
        if op == ">":
            return a > b
        if op == "<":
            return a < b
        if op == ">=":
            return a >= b
        if op == "<=":
            return a <= b
        if op == "==":
            return a == b
        if op == "!=":
            return a != b
        if op == "and":
            return np.bitwise_and(a, b)
        if op == "or":
            return np.bitwise_or(a, b)
        raise NotImplementedError(op)
    if t == "ternary":
        return np.where(_eval(node["cond"], ctx, reg), _eval(node["true"], ctx, reg), _eval(node["false"], ctx, reg))
    if t == "call":
        fn = _NATIVE[node["func"]]
        args = [_eval(a, ctx, reg) for a in node["args"]]
        return fn(*args)
    raise NotImplementedError(t)


def _local_key(ctx, reg):
    # locals are allocated over the current block region; index relative to it
    i0, i1, j0, j1, k0, k1 = reg
    bi0, bj0, bk0 = ctx.local_base
    return (slice(i0 - bi0, i1 - bi0), slice(j0 - bj0, j1 - bj0), slice(k0 - bk0, k1 - bk0))


def _bound_abs(b, n, default):
    if b is None:
        return default
    level, off = b
    return off if level == "start" else n + off


class _LazyMask:
    """A mask that is RE-EVALUATED by every statement it guards: how the numpy backend lowers `while` — the loop
    condition (AND the enclosing masks) is inlined as the `np.where` condition of every statement of the body
    (gtc/numpy/oir_to_npir.py:149-185: visit_While passes `mask=cond_expr` down, visit_MaskStmt ANDs onto it,
    visit_AssignStmt wraps the right-hand side in a VectorTernaryOp on that expression)."""

    def __init__(self, fn):
        self.fn = fn

    def __call__(self, ctx, reg):
        return self.fn(ctx, reg)


def _mask_now(mask, ctx, reg):
    return mask(ctx, reg) if isinstance(mask, _LazyMask) else mask


def _assign(stmt, ctx, reg, mask):
    left = stmt["left"]
    i0, i1, j0, j1, k0, k1 = reg
    if i1 <= i0 or j1 <= j0 or k1 <= k0:
        return
    mask = _mask_now(mask, ctx, reg)
    rhs = _eval(stmt["right"], ctx, reg)
    name = left["name"]
    if name in ctx.locals:
        tgt = ctx.locals[name]
        key = _local_key(ctx, reg)
        if mask is not None:
            rhs = np.where(mask, rhs, tgt[key])
        tgt[key] = rhs
        return
    f = ctx.fields[name]
    off = left["off"]
    data_index = tuple(int(_eval(d, ctx, reg)) for d in left.get("data_index", []))
    if isinstance(off, dict):
        if "vk" in off:
            dk = _eval(off["vk"], ctx, reg)
            kidx = np.arange(k0, k1)[None, None, :] + dk
            kidx = np.broadcast_to(kidx, (i1 - i0, j1 - j0, k1 - k0)) if np.ndim(kidx) == 3 and kidx.shape[0] == 1 else kidx
        else:
            raise NotImplementedError("absolute K write")
        ksl = kidx
    else:
        ksl = slice(k0 + off[2], k1 + off[2])
    if mask is not None:
        cur = f.get(slice(i0, i1), slice(j0, j1), ksl, data_index)
        rhs = np.where(mask, rhs, cur)
    f.set(slice(i0, i1), slice(j0, j1), ksl, data_index, rhs)


def _exec_stmts(stmts, ctx, reg, mask):
    for s in stmts:
        t = s["t"]
        if t == "assign":
            _assign(s, ctx, reg, mask)
        elif t == "mask":
            if isinstance(mask, _LazyMask):  # inside a while body: the whole mask chain is re-evaluated per statement
                m = _LazyMask(lambda c, r, outer=mask, e=s["mask"]: np.bitwise_and(outer(c, r), _eval(e, c, r)))
            else:
                m = _eval(s["mask"], ctx, reg)
                if mask is not None:
                    m = np.bitwise_and(mask, m)
            _exec_stmts(s["body"], ctx, reg, m)
        elif t == "while":
            # numpy-backend semantics (north_star's oracle): `while np.any(cond): <every body statement masked by the
            # re-evaluated cond>` (npir_codegen.py:252-267).  Equal to a per-point `while` whenever the loop variable is
            # updated by the last statement of the body; fixture while_first_f64 pins the case where it is not.
            if mask is None:
                active = _LazyMask(lambda c, r, e=s["cond"]: _eval(e, c, r))
            else:
                active = _LazyMask(lambda c, r, outer=mask, e=s["cond"]: np.bitwise_and(_mask_now(outer, c, r), _eval(e, c, r)))
            while np.any(active(ctx, reg)):
                _exec_stmts(s["body"], ctx, reg, active)
        elif t == "hregion":
            i0, i1, j0, j1, k0, k1 = reg
            ri0 = max(i0, _bound_abs(s["i"][0], ctx.nI, -(10**9)))
            ri1 = min(i1, _bound_abs(s["i"][1], ctx.nI, 10**9))
            rj0 = max(j0, _bound_abs(s["j"][0], ctx.nJ, -(10**9)))
            rj1 = min(j1, _bound_abs(s["j"][1], ctx.nJ, 10**9))
            if ri1 <= ri0 or rj1 <= rj0:
                continue
            sub = (ri0, ri1, rj0, rj1, k0, k1)
            if isinstance(mask, _LazyMask):
                raise NotImplementedError("horizontal region inside a while loop")
            m = mask
            if m is not None and isinstance(m, np.ndarray) and m.ndim == 3:
                m = m[
                    (slice(ri0 - i0, ri1 - i0) if m.shape[0] != 1 else slice(None)),
                    (slice(rj0 - j0, rj1 - j0) if m.shape[1] != 1 else slice(None)),
                    slice(None),
                ]
            _exec_stmts(s["body"], ctx, sub, m)
        else:
            raise NotImplementedError(t)


def _alloc_locals(he, ctx, reg):
    i0, i1, j0, j1, k0, k1 = reg
    ctx.locals = {
        d["name"]: np.zeros((i1 - i0, j1 - j0, k1 - k0), dtype=_NP_DTYPE[d["dtype"]]) for d in he["locals"]
    }
    ctx.local_base = (i0, j0, k0)


def run(stencil: Dict[str, Any], fields: Dict[str, np.ndarray], params: Dict[str, Any],
        domain: Tuple[int, int, int], origins: Dict[str, Tuple[int, ...]]) -> None:
    """Execute `stencil` in place on host arrays (arrays are in IJK[+data] axis order)."""
    nI, nJ, nK = (int(d) for d in domain)
    fobjs: Dict[str, _Field] = {}
    for p in stencil["params"]:
        if p["t"] == "field" and fields.get(p["name"]) is not None:
            fobjs[p["name"]] = _Field(fields[p["name"]], origins[p["name"]], p["dims"])
    for tmp in stencil["temporaries"]:
        (ei0, ei1), (ej0, ej1) = tmp["extent"]
        shape = [nI + (ei1 - ei0), nJ + (ej1 - ej0)]
        origin = [-ei0, -ej0]
        if tmp["dims"][2]:
            shape.append(nK)
            origin.append(0)
        shape += list(tmp["data_dims"])
        origin += [0] * len(tmp["data_dims"])
        dims = [True, True, bool(tmp["dims"][2])]
        fobjs[tmp["name"]] = _Field(np.zeros(shape, dtype=_NP_DTYPE[tmp["dtype"]]), tuple(origin), dims)
    cparams = {}
    for p in stencil["params"]:
        if p["t"] == "scalar" and p["name"] in params and params[p["name"]] is not None:
            cparams[p["name"]] = _NP_DTYPE[p["dtype"]](params[p["name"]])
    ctx = _Ctx(fobjs, cparams, (nI, nJ, nK))

    with np.errstate(divide="ignore", over="ignore", under="ignore", invalid="ignore"):
        for loop in stencil["loops"]:
            order = loop["order"]
            for sec in loop["sections"]:
                k0 = _bound_abs(sec["interval"][0], nK, 0)
                k1 = _bound_abs(sec["interval"][1], nK, nK)
                if order == "parallel":
                    levels = [(k0, k1)]
                elif order == "forward":
                    levels = [(k, k + 1) for k in range(k0, k1)]
                else:
                    levels = [(k, k + 1) for k in range(k1 - 1, k0 - 1, -1)]
                for ka, kb in levels:
                    for he in sec["hes"]:
                        (ei0, ei1), (ej0, ej1) = he["extent"]
                        reg = (ei0, nI + ei1, ej0, nJ + ej1, ka, kb)
                        _alloc_locals(he, ctx, reg)
                        _exec_stmts(he["body"], ctx, reg, None)
    ctx.locals = {}


def default_origins(stencil: Dict[str, Any]) -> Dict[str, Tuple[int, ...]]:
    """Smallest legal origin of every API field (its lower boundary)."""
    out = {}
    for name, fi in stencil["field_info"].items():
        if fi is None:
            continue
        axes = fi["axes"]
        b = fi["boundary"]
        o = [b["IJK".index(a)][0] for a in axes] + [0] * len(fi["data_dims"])
        out[name] = tuple(o)
    return out
