"""b200 stencil IR -> sm_100a CUDA source + launch plan.

This is the backend code generator of the `b200` entry: the component that stands in for the
reference's GridTools C++ generator (reference: gtc/gtcpp/gtcpp_codegen.py:37-331,
backend/gtcpp_backend.py:35-62) — written from scratch, emitting plain SIMT CUDA kernels, one per
`computation` block (or per group of horizontal executions that can run without a grid-wide
barrier), instead of GridTools template expressions.

Two kernel families are emitted:

* ``par``  PARALLEL vertical loops: one thread per (i, j) point, K from ``blockIdx.z``; the loop's
  sections become branches on ``k``.  Horizontal executions are fused into one kernel as long as no
  field written by an earlier execution is read at a horizontal offset by a later one (that needs a
  plane-wide barrier -> next kernel); surviving temporaries live in backend-owned scratch.
* ``seq``  FORWARD/BACKWARD vertical loops: one thread per (i, j) column, the K loop stays inside the
  thread and sections become consecutive sub-loops.  Loops whose temporaries are read at horizontal
  offsets in the same level fall back to level-by-level ``par`` launches.

The tiled/streaming generator for multi-stage PARALLEL blocks lives in `codegen_stream.py` and is
tried first by `generate()`; everything it cannot handle comes here.

Semantics follow the `numpy` backend (the oracle), see SURVEY.md §9 / oracle/numpy_oracle.py.
"""

from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

from . import ir as b2ir

CT = b2ir.CTYPE

_FUNC = {
    "abs": "abs_", "min": "min_", "max": "max_", "mod": "mod_", "sin": "sin_", "cos": "cos_",
    "tan": "tan_", "arcsin": "asin_", "arccos": "acos_", "arctan": "atan_", "sinh": "sinh_",
    "cosh": "cosh_", "tanh": "tanh_", "arcsinh": "asinh_", "arccosh": "acosh_", "arctanh": "atanh_",
    "sqrt": "sqrt_", "pow": "pow_", "exp": "exp_", "log": "log_", "log10": "log10_", "gamma": "gamma_",
    "cbrt": "cbrt_", "isfinite": "isfinite_", "isinf": "isinf_", "isnan": "isnan_", "floor": "floor_",
    "ceil": "ceil_", "trunc": "trunc_", "erf": "erf_", "erfc": "erfc_", "round": "round_",
    "round_away_from_zero": "round_away_",
}  # fmt: skip
_CASTFUNC = {"int32": "int32", "int64": "int64", "float32": "float32", "float64": "float64"}


class CodegenError(NotImplementedError):
    pass


def literal(value: str, dtype: str) -> str:
    if dtype == "bool":
        return "true" if value in ("True", "true", "1") else "false"
    if dtype.startswith("int"):
        iv = int(value)
        if dtype == "int64":
            return f"({iv}LL)" if iv > -(2**63) else "(-9223372036854775807LL-1)"
        return f"(({CT[dtype]}){iv})"
    fv = float(value)
    if fv != fv:
        return "(__int_as_float(0x7fc00000))" if dtype == "float32" else "(__longlong_as_double(0x7ff8000000000000LL))"
    if fv in (float("inf"), float("-inf")):
        base = "__int_as_float(0x7f800000)" if dtype == "float32" else "__longlong_as_double(0x7ff0000000000000LL)"
        return f"({'-' if fv < 0 else ''}{base})"
    if dtype == "float32":
        import numpy as np

        return f"({float(np.float32(fv)).hex()}f)"
    return f"({fv.hex()})"


class FieldTable:
    """Index of API fields + temporaries in the kernel argument block."""

    def __init__(self, stencil):
        self.stencil = stencil
        self.entries: List[dict] = []
        self.index: Dict[str, int] = {}
        for p in stencil["params"]:
            if p["t"] == "field":
                self._add(p["name"], p["dtype"], p["dims"], p["data_dims"], "api", None)
        for t in stencil["temporaries"]:
            self._add(t["name"], t["dtype"], t["dims"], t["data_dims"], "temp", t["extent"])
        self.scalars = [p for p in stencil["params"] if p["t"] == "scalar"]

    def _add(self, name, dtype, dims, data_dims, kind, extent):
        if len(data_dims) > 4:
            raise CodegenError("b200: more than four data dimensions")
        self.index[name] = len(self.entries)
        self.entries.append(
            {"name": name, "dtype": dtype, "dims": list(dims), "data_dims": list(data_dims), "kind": kind, "extent": extent}
        )

    def scalar_layout(self) -> Tuple[List[dict], int]:
        """C layout of the scalar-parameter blob (natural alignment, API order)."""
        off = 0
        out = []
        for p in self.scalars:
            size = b2ir.ITEMSIZE[p["dtype"]]
            off = (off + size - 1) // size * size
            out.append({"name": p["name"], "dtype": p["dtype"], "offset": off})
            off += size
        return out, (off + 7) // 8 * 8


class ExprGen:
    """Expression / statement emitter for point-wise (one thread = one (i,j[,k]) point) code."""

    def __init__(self, ft: FieldTable, written: set, *, ivar="i", jvar="j", kvar="k", args="A"):
        self.ft = ft
        self.written = written  # fields written by the current kernel (no read-only loads)
        self.i, self.j, self.k, self.A = ivar, jvar, kvar, args
        self.locals: Dict[str, str] = {}
        #: callable(C expression of a launch-invariant divisor) -> name of its hoisted b200::DivInv, or None: plain division
        self.div_hoist = None
        #: callable(C expression of the dividend, DivInv name, C type) -> name of the quotient variable, or None: b200::div_inv
        self.div_group = None

    def invariant(self, n) -> bool:
        """The expression has the same value for every cell of a launch: literals and scalar parameters only."""
        t = n["t"]
        if t == "lit":
            return True
        if t == "scalar":
            return n["name"] not in self.locals
        if t in ("cast", "unary"):
            return self.invariant(n["expr"])
        if t == "binary":
            return self.invariant(n["left"]) and self.invariant(n["right"])
        if t == "ternary":
            return self.invariant(n["cond"]) and self.invariant(n["true"]) and self.invariant(n["false"])
        if t == "call":
            return all(self.invariant(a) for a in n["args"])
        return False

    # -- addressing ---------------------------------------------------------------------------
    def field_ref(self, node, *, for_write=False) -> str:
        name = node["name"]
        n = self.ft.index[name]
        ent = self.ft.entries[n]
        ct = CT[ent["dtype"]]
        f = f"{self.A}.f[{n}]"
        off = node["off"]
        terms = []
        if isinstance(off, dict):
            di = dj = 0
            if "vk" in off:
                kexpr = f"b200::clampk((long long){self.k} + (long long)({self.expr(off['vk'])}), {f}.klo, {f}.khi)"
            else:
                ak = off["abs_k"]
                kexpr = f"((long long)({ak if isinstance(ak, int) else self.expr(ak)}))"
        else:
            di, dj, dk = off
            kexpr = f"({self.k}{dk:+d})" if dk else self.k
        if ent["dims"][0]:
            terms.append(f"(long long)({self.i}{di:+d})*{f}.s[0]" if di else f"(long long){self.i}*{f}.s[0]")
        if ent["dims"][1]:
            terms.append(f"(long long)({self.j}{dj:+d})*{f}.s[1]" if dj else f"(long long){self.j}*{f}.s[1]")
        if ent["dims"][2]:
            terms.append(f"(long long){kexpr}*{f}.s[2]")
        for d, ix in enumerate(node.get("data_index", [])):
            terms.append(f"(long long)({self.expr(ix)})*{f}.s[{3 + d}]")
        idx = " + ".join(terms) if terms else "0"
        return f"(({ct}*){f}.p)[{idx}]"

    def field_load(self, node) -> str:
        ref = self.field_ref(node)
        if node["name"] in self.written:
            return ref
        ct = CT[self.ft.entries[self.ft.index[node["name"]]]["dtype"]]
        return f"b200::ldro<{ct}>(&{ref})"

    # -- expressions ----------------------------------------------------------------------------
    def expr(self, n) -> str:
        t = n["t"]
        if t == "field":
            return self.field_load(n)
        if t == "scalar":
            name = n["name"]
            if name in self.locals:
                return self.locals[name]
            return f"{self.A}.p_{name}"
        if t == "lit":
            return literal(n["value"], n["dtype"])
        if t == "iter":
            axis = {"I": self.i, "J": self.j, "K": self.k}[n["axis"]]
            return f"(({CT[n['dtype']]}){axis})"
        if t == "cast":
            src = n["expr"]
            if n["dtype"] == "bool" and src["dtype"] != "bool":
                return f"(({self.expr(src)}) != 0)"
            return f"(({CT[n['dtype']]})({self.expr(src)}))"
        if t == "unary":
            op = n["op"]
            e = self.expr(n["expr"])
            if op == "not":
                return f"(!({e}))"
            if op == "-":
                return f"(({CT[n['dtype']]})(-({e})))"
            return f"(+({e}))"
        if t == "binary":
            op = n["op"]
            a, b = self.expr(n["left"]), self.expr(n["right"])
            if op in ("and", "or"):
                return f"(({a}) {'&&' if op == 'and' else '||'} ({b}))"
            if op in (">", "<", ">=", "<=", "==", "!="):
                return f"(({a}) {op} ({b}))"
            dt = n["dtype"]
            if dt in ("int8", "int16", "bool"):
                return f"(({CT[dt]})(({a}) {op} ({b})))"
            if op == "/" and dt.startswith("int"):
                # NumPy true-divide then store into an integer: truncation toward zero == C division
                return f"((({b}) == 0) ? ({CT[dt]})0 : ({CT[dt]})(({a}) / ({b})))"
            if op == "/" and dt in ("float32", "float64") and self.div_hoist is not None and self.invariant(n["right"]):
                # the divisor is the same for every cell of the launch: hoisted with its reciprocal, 3 FP
                # instructions per cell, bit-equal to the IEEE division (b200_device.cuh, DivInv)
                dv = self.div_hoist(f"({CT[dt]})({b})")
                if self.div_group is not None:  # the emitter guards a whole group of quotients with one branch
                    return self.div_group(f"({CT[dt]})({a})", dv, CT[dt])
                return f"b200::div_inv(({CT[dt]})({a}), {dv})"
            return f"(({a}) {op} ({b}))"
        if t == "ternary":
            ct = CT[n["dtype"]]
            return f"(({self.expr(n['cond'])}) ? ({ct})({self.expr(n['true'])}) : ({ct})({self.expr(n['false'])}))"
        if t == "call":
            fn = n["func"]
            args = [self.expr(a) for a in n["args"]]
            if fn in _CASTFUNC:
                return f"(({CT[_CASTFUNC[fn]]})({args[0]}))"
            if fn not in _FUNC:
                raise CodegenError(f"b200: native function {fn}")
            if fn == "pow" and n["dtype"].startswith("float"):
                # the upcaster leaves pow's operands alone (gtc/passes/gtir_upcaster.py:88-143):
                # np.power(int, float) is a floating power, never an integer one
                args = [f"(({CT[n['dtype']]})({a}))" for a in args]
            call = f"b200::{_FUNC[fn]}({', '.join(args)})"
            if fn in ("isfinite", "isinf", "isnan"):
                return call
            return f"(({CT[n['dtype']]}){call})"
        raise CodegenError(f"b200: expression {t}")

    # -- statements -----------------------------------------------------------------------------
    def stmts(self, body, ind: str, masks: Tuple[str, ...] = (), lazy: bool = False) -> List[str]:
        """Statements of one point.  `masks`: the C conditions of the enclosing `if` / `while` statements.

        Outside a `while` a mask is evaluated once (an `if` block).  INSIDE a `while` (lazy=True) every assignment is
        guarded by the whole chain, re-evaluated at that statement: the numpy backend — north_star's oracle — inlines
        the loop condition (AND the enclosing masks) as the `np.where` condition of every statement of the body
        (gtc/numpy/oir_to_npir.py:149-185, npir_codegen.py:252-267), which differs from a per-point `while` when the
        loop variable is not updated by the last statement (fixture while_first_f64)."""
        out: List[str] = []
        for s in body:
            t = s["t"]
            if t == "assign":
                left = s["left"]
                rhs = self.expr(s["right"])
                if left["t"] == "scalar":
                    line = f"{self.locals[left['name']]} = ({CT[left['dtype']]})({rhs});"
                else:
                    ct = CT[self.ft.entries[self.ft.index[left['name']]]["dtype"]]
                    line = f"{self.field_ref(left, for_write=True)} = ({ct})({rhs});"
                out.append(f"{ind}if ({' && '.join(masks)}) {{ {line} }}" if lazy and masks else f"{ind}{line}")
            elif t in ("mask", "hregion"):
                if t == "mask":
                    cond = f"({self.expr(s['mask'])})"
                else:
                    conds = []
                    for var, n_sym, (lo, hi) in ((self.i, f"{self.A}.g.nI", s["i"]), (self.j, f"{self.A}.g.nJ", s["j"])):
                        if lo is not None:
                            conds.append(f"{var} >= {_bound(lo, n_sym)}")
                        if hi is not None:
                            conds.append(f"{var} < {_bound(hi, n_sym)}")
                    cond = f"({' && '.join(conds) if conds else 'true'})"
                if lazy:
                    out += self.stmts(s["body"], ind, masks + (cond,), True)
                else:
                    out.append(f"{ind}if {cond} {{")
                    out += self.stmts(s["body"], ind + "  ", masks + (cond,), False)
                    out.append(f"{ind}}}")
            elif t == "while":
                chain = masks + (f"({self.expr(s['cond'])})",)
                out.append(f"{ind}while ({' && '.join(chain)}) {{")
                out += self.stmts(s["body"], ind + "  ", chain, True)
                out.append(f"{ind}}}")
            else:
                raise CodegenError(f"b200: statement {t}")
        return out

    def he_block(self, he, ind: str, *, guard_extent=True, declare_locals=True) -> List[str]:
        """One horizontal execution for the current point, guarded by its own block extent."""
        (ei0, ei1), (ej0, ej1) = he["extent"]
        out = []
        A = self.A
        if guard_extent:
            out.append(
                f"{ind}if ({self.i} >= {A}.g.i_lo{ei0:+d} && {self.i} < {A}.g.i_hi{ei1:+d} && "
                f"{self.j} >= {A}.g.j_lo{ej0:+d} && {self.j} < {A}.g.j_hi{ej1:+d}) {{"
            )
        else:
            out.append(f"{ind}{{")
        saved = dict(self.locals)
        if declare_locals:
            for d in he["locals"]:
                self.locals[d["name"]] = f"l_{d['name']}"
                out.append(f"{ind}  {CT[d['dtype']]} l_{d['name']} = ({CT[d['dtype']]})0;")
        out += self.stmts(he["body"], ind + "  ")
        out.append(f"{ind}}}")
        if declare_locals:
            self.locals = saved
        return out


def _bound(b, n_sym: str) -> str:
    level, off = b
    return f"({off})" if level == "start" else f"({n_sym}{off:+d})"


def _written_in(hes) -> set:
    return {a["name"] for he in hes for a in b2ir.field_accesses(he["body"]) if a["write"]}


def split_groups(hes: List[dict]) -> List[List[dict]]:
    """Split a section's horizontal executions where a plane-wide barrier is needed.

    A new kernel starts when an execution reads, at a non-zero IJ offset, a field written earlier in
    the group (RAW across threads) or writes a field read earlier at a non-zero IJ offset (WAR).
    """
    groups: List[List[dict]] = []
    cur: List[dict] = []
    written: set = set()
    read_off: set = set()
    for he in hes:
        acc = b2ir.field_accesses(he["body"])
        raw = any((not a["write"]) and a["name"] in written and b2ir.ij_offset(a["off"]) != (0, 0) for a in acc)
        war = any(a["write"] and a["name"] in read_off for a in acc)
        if cur and (raw or war):
            groups.append(cur)
            cur, written, read_off = [], set(), set()
        cur.append(he)
        written |= {a["name"] for a in acc if a["write"]}
        read_off |= {a["name"] for a in acc if not a["write"] and b2ir.ij_offset(a["off"]) != (0, 0)}
    if cur:
        groups.append(cur)
    return groups


def _union_extent(hes) -> List[List[int]]:
    e = [[0, 0], [0, 0]]
    for he in hes:
        for a in range(2):
            e[a][0] = min(e[a][0], he["extent"][a][0])
            e[a][1] = max(e[a][1], he["extent"][a][1])
    return e


def _sections_k_coupled(loop) -> bool:
    """A section of this PARALLEL loop reads, at a K offset (or a variable / absolute K index), a field that ANOTHER
    section of the loop writes.  The reference merges adjacent-interval PARALLEL loops without a dependency check
    (gtc/passes/oir_optimizations/vertical_loop_merging.py, AdjacentLoopMerging) and its numpy backend runs the sections
    one after the other, so `interval(0,1): b = a; interval(1,None): c = b[0,0,-1]` is legal and ordered: the sections
    must then be separate launches, in order (one kernel with K on blockIdx.z would race across levels)."""
    secs = loop["sections"]
    if len(secs) < 2:
        return False
    writes, koff_reads = [], []
    for sec in secs:
        acc = [a for he in sec["hes"] for a in b2ir.field_accesses(he["body"])]
        writes.append({a["name"] for a in acc if a["write"]})
        koff_reads.append({a["name"] for a in acc if not a["write"] and (isinstance(a["off"], dict) or a["off"][2] != 0)})
    for n, reads in enumerate(koff_reads):
        for m, w in enumerate(writes):
            if m != n and reads & w:
                return True
    return False


def _needs_level_sync(loop) -> bool:
    written, read_off = set(), set()
    for sec in loop["sections"]:
        for he in sec["hes"]:
            for a in b2ir.field_accesses(he["body"]):
                if a["write"]:
                    written.add(a["name"])
                elif b2ir.ij_offset(a["off"]) != (0, 0):
                    read_off.add(a["name"])
    return bool(written & read_off)


class Generator:
    """Baseline ("point") generator: always applicable."""

    BLOCK_PAR = (64, 4)
    BLOCK_SEQ = (64, 2)

    def __init__(self, stencil: Dict[str, Any], options: Optional[Dict[str, Any]] = None):
        self.st = stencil
        self.opt = dict(options or {})
        self.ft = FieldTable(stencil)
        self.kernels: List[dict] = []
        self.steps: List[dict] = []
        self.src: List[str] = []
        self.live: set = set()  # fields some kernel addresses in global memory
        self.tmaps: List[Dict[str, Any]] = []  # tensor maps the launcher encodes into the argument block

    # -- argument block ---------------------------------------------------------------------------
    def args_struct(self) -> str:
        scal, size = self.ft.scalar_layout()
        lines = ["struct Args {", "  b200::Geom g;", f"  b200::FieldArg f[{max(1, len(self.ft.entries))}];"]
        pos = 0
        pad = 0
        for s in scal:
            if s["offset"] > pos:
                lines.append(f"  char _pad{pad}[{s['offset'] - pos}];")
                pad += 1
            lines.append(f"  {CT[s['dtype']]} p_{s['name']};")
            pos = s["offset"] + b2ir.ITEMSIZE[s["dtype"]]
        if size > pos:
            lines.append(f"  char _pad{pad}[{size - pos}];")
        if self.tmaps:
            # tensor maps of the bulk-async streaming kernels, encoded by the launcher at every call: 64-byte aligned
            # right behind the scalars; tmo = {array index of the domain origin along I (incl. the alignment shift of
            # the map's base), J, K, K multiplier (0 for a field without K axis)} per map
            lines.append(f"  b200::TMap tm[{len(self.tmaps)}];")
            lines.append(f"  int tmo[{len(self.tmaps)}][4];")
        lines.append("};")
        return "\n".join(lines)

    def tmap_index(self, field: str, box0: int, box1: int) -> int:
        """Index (in Args::tm) of the tensor map of `field` with a box of box0 elements x box1 rows."""
        fi = self.ft.index[field]
        for n, t in enumerate(self.tmaps):
            if (t["field"], t["box"]) == (fi, [box0, box1]):
                return n
        self.tmaps.append({"field": fi, "box": [box0, box1]})
        return len(self.tmaps) - 1

    # -- kernels ------------------------------------------------------------------------------------
    def _kname(self, tag: str) -> str:
        return f"b200_{_cname(self.st['name'])}_{tag}{len(self.kernels)}"

    def emit_par_kernel(self, sections: List[Tuple[list, List[dict]]], *, level_mode=False) -> int:
        """sections: [(interval, [he...])]; all fused in one kernel, K from blockIdx.z."""
        name = self._kname("par")
        all_hes = [he for _, hes in sections for he in hes]
        self.live |= {a["name"] for he in all_hes for a in b2ir.field_accesses(he["body"])}
        eg = ExprGen(self.ft, _written_in(all_hes))
        ext = _union_extent(all_hes)
        bx, by = self.BLOCK_PAR
        L = [f'extern "C" __global__ void __launch_bounds__({bx * by}) {name}(const __grid_constant__ Args A) {{']
        L.append(f"  const int i = A.g.i_lo + ({ext[0][0]}) + (int)(blockIdx.x * {bx} + threadIdx.x);")
        L.append(f"  const int j = A.g.j_lo + ({ext[1][0]}) + (int)(blockIdx.y * {by} + threadIdx.y);")
        L.append("  const int k = A.g.k_lo + (int)blockIdx.z;")
        for interval, hes in sections:
            k0 = _bound(interval[0], "A.g.nK")
            k1 = _bound(interval[1], "A.g.nK")
            L.append(f"  if (k >= {k0} && k < {k1}) {{")
            for he in hes:
                L += eg.he_block(he, "    ")
            L.append("  }")
        L.append("}")
        self.src.append("\n".join(L))
        k_lo = sections[0][0][0]
        k_hi = sections[-1][0][1]
        self.kernels.append(
            {"name": name, "kind": "par", "block": [bx, by, 1], "extent": ext, "k_lo": k_lo, "k_hi": k_hi, "smem": 0}
        )
        return len(self.kernels) - 1

    def emit_seq_kernel(self, loop) -> int:
        name = self._kname("seq")
        order = loop["order"]
        all_hes = [he for sec in loop["sections"] for he in sec["hes"]]
        self.live |= {a["name"] for he in all_hes for a in b2ir.field_accesses(he["body"])}
        eg = ExprGen(self.ft, _written_in(all_hes))
        ext = _union_extent(all_hes)
        bx, by = self.BLOCK_SEQ
        L = [f'extern "C" __global__ void __launch_bounds__({bx * by}) {name}(const __grid_constant__ Args A) {{']
        L.append(f"  const int i = A.g.i_lo + ({ext[0][0]}) + (int)(blockIdx.x * {bx} + threadIdx.x);")
        L.append(f"  const int j = A.g.j_lo + ({ext[1][0]}) + (int)(blockIdx.y * {by} + threadIdx.y);")
        for sec in loop["sections"]:
            k0 = _bound(sec["interval"][0], "A.g.nK")
            k1 = _bound(sec["interval"][1], "A.g.nK")
            if order == "forward":
                L.append(f"  for (int k = {k0}; k < {k1}; ++k) {{")
            else:
                L.append(f"  for (int k = {k1} - 1; k >= {k0}; --k) {{")
            for he in sec["hes"]:
                L += eg.he_block(he, "    ")
            L.append("  }")
        L.append("}")
        self.src.append("\n".join(L))
        self.kernels.append(
            {"name": name, "kind": "seq", "block": [bx, by, 1], "extent": ext, "k_lo": ["start", 0], "k_hi": ["start", 1], "smem": 0}
        )
        return len(self.kernels) - 1

    # -- driver -------------------------------------------------------------------------------------
    # (one kernel for all sections of a PARALLEL loop takes K from blockIdx.z: the sections then run concurrently)
    def lower_loop(self, loop) -> None:
        order = loop["order"]
        if order == "parallel":
            grouped = [(sec["interval"], split_groups(sec["hes"])) for sec in loop["sections"]]
            if all(len(g) == 1 for _, g in grouped) and not _sections_k_coupled(loop):
                k = self.emit_par_kernel([(iv, g[0]) for iv, g in grouped])
                self.steps.append({"t": "launch", "kernel": k})
            else:
                for iv, groups in grouped:
                    for g in groups:
                        k = self.emit_par_kernel([(iv, g)])
                        self.steps.append({"t": "launch", "kernel": k})
            return
        if not _needs_level_sync(loop):
            from . import codegen_column

            k = codegen_column.try_emit(self, loop, self.opt)  # register k-caches + prefetch
            if k is None:
                k = self.emit_seq_kernel(loop)
            self.steps.append({"t": "launch", "kernel": k})
            return
        # level-by-level fallback: sections in loop order, every group its own kernel
        secs = []
        for sec in loop["sections"]:
            ks = [self.emit_par_kernel([(sec["interval"], g)], level_mode=True) for g in split_groups(sec["hes"])]
            secs.append({"interval": sec["interval"], "kernels": ks})
        self.steps.append({"t": "levels", "order": order, "sections": secs})

    def lower_loops(self, loops: List[dict]) -> None:
        """Lower a run of consecutive loops; consecutive FORWARD/BACKWARD sweeps whose data flow stays
        inside a column become ONE column kernel (forward elimination + back substitution of a
        tridiagonal solve in one launch) when `fuse_columns` is on.  Off by default: the fused kernel
        runs every sweep at the register count of the heaviest one (Thomas: the 32-register back
        substitution would run at the 58 registers of the elimination, 36 instead of 64 warps/SM), which
        can cost more than the saved launch and the L2 hits on the freshest levels — to be measured."""
        from . import codegen_column

        n = 0
        while n < len(loops):
            loop = loops[n]
            group = [loop]
            if loop["order"] != "parallel" and not _needs_level_sync(loop) and (self.opt.get("fuse_columns", False) or self.opt.get("col_smem", False)):
                m = n + 1
                while (m < len(loops) and loops[m]["order"] != "parallel" and not _needs_level_sync(loops[m])
                       and all(codegen_column.fusable(g, loops[m]) for g in group)):  # fmt: skip
                    group.append(loops[m])
                    m += 1
            if len(group) > 1:
                # fields of the loops outside this kernel: temporaries none of them touches may live in shared memory
                external = {a["name"] for lp in self.st["loops"] if not any(lp is g for g in group)
                            for sec in lp["sections"] for he in sec["hes"] for a in b2ir.field_accesses(he["body"])}  # fmt: skip
                k = codegen_column.try_emit(self, group, self.opt, external)
                if k is not None and not self.opt.get("fuse_columns", False) and not self.kernels[k].get("smem_fields"):
                    # fused only for the sake of shared-memory temporaries and none qualified: keep the separate sweeps
                    self.kernels.pop()
                    self.src.pop()
                    k = None
                if k is not None:
                    self.steps.append({"t": "launch", "kernel": k})
                    n += len(group)
                    continue
            self.lower_loop(loop)
            n += 1

    def generate(self) -> Tuple[str, Dict[str, Any]]:
        self.lower_loops(self.st["loops"])
        return self.finish()

    def finish(self) -> Tuple[str, Dict[str, Any]]:
        scal, scal_size = self.ft.scalar_layout()
        header = [
            "// generated by gt4py_b200.codegen — do not edit",
            f"// stencil: {self.st['name']}",
            '#include "b200_device.cuh"',
            "namespace {",
            self.args_struct(),
            "}  // namespace",
            "",
        ]
        source = "\n".join(header) + "\n\n".join(self.src) + "\n"
        plan = {
            "version": 3,
            "name": self.st["name"],
            "fields": [
                {
                    "name": e["name"],
                    "dtype": e["dtype"],
                    "itemsize": b2ir.ITEMSIZE[e["dtype"]],
                    "dims": e["dims"],
                    "data_dims": e["data_dims"],
                    "kind": e["kind"] if (e["kind"] == "api" or e["name"] in self.live) else "dead",
                    "extent": e["extent"],
                }
                for e in self.ft.entries
            ],
            "scalars": scal,
            "scalars_size": scal_size,
            "kernels": self.kernels,
            "steps": self.steps,
            "tmaps": self.tmaps,
        }
        return source, plan


def _cname(name: str) -> str:
    return "".join(c if c.isalnum() else "_" for c in name)


def generate(stencil: Dict[str, Any], options: Optional[Dict[str, Any]] = None) -> Tuple[str, Dict[str, Any]]:
    """IR -> (CUDA source, launch plan).  Tries the streaming generator first (if enabled)."""
    options = dict(options or {})
    if options.get("strategy", "auto") in ("auto", "stream"):
        try:
            from . import codegen_stream

            res = codegen_stream.try_generate(stencil, options)
            if res is not None:
                return res
        except ImportError:
            pass
        if options.get("strategy") == "stream":
            raise CodegenError("b200: streaming strategy not applicable to this stencil")
    return Generator(stencil, options).generate()


def plan_to_text(plan: Dict[str, Any]) -> str:
    """Flat, line-oriented rendering of the launch plan for the C launcher (csrc/launcher.cu)."""

    def b(bound):
        return f"{0 if bound[0] == 'start' else 1} {bound[1]}"

    L = [f"b200plan {plan['version']}", f"name {plan['name']}"]
    L.append(f"nfields {len(plan['fields'])}")
    for f in plan["fields"]:
        e = f["extent"] or [[0, 0], [0, 0]]
        dd = f["data_dims"] + [1] * (4 - len(f["data_dims"]))
        L.append(
            f"field {f['name']} { {'api': 0, 'temp': 1, 'dead': 2}[f['kind']] } {f['itemsize']} "
            f"{int(f['dims'][0])} {int(f['dims'][1])} {int(f['dims'][2])} {len(f['data_dims'])} {dd[0]} {dd[1]} {dd[2]} {dd[3]} "
            f"{e[0][0]} {e[0][1]} {e[1][0]} {e[1][1]}"
        )
    L.append(f"scalars_size {plan['scalars_size']}")
    L.append(f"nkernels {len(plan['kernels'])}")
    for k in plan["kernels"]:
        e = k["extent"]
        kind = {"par": 0, "seq": 1, "stream": 2}[k["kind"]]
        tile = k.get("tile", [k["block"][0], k["block"][1], 1])
        L.append(
            f"kernel {k['name']} {kind} {k['block'][0]} {k['block'][1]} {k['block'][2]} "
            f"{tile[0]} {tile[1]} {tile[2]} {e[0][0]} {e[0][1]} {e[1][0]} {e[1][1]} {b(k['k_lo'])} {b(k['k_hi'])} {k['smem']} {int(k.get('qshift', 0))} "
            f"{int(k.get('smem_per_k', 0))} {int(k.get('smem_kcap', 0))}"
        )
    L.append(f"ntmaps {len(plan.get('tmaps', []))}")
    for t in plan.get("tmaps", []):
        L.append(f"tmap {t['field']} {t['box'][0]} {t['box'][1]}")
    L.append(f"nsteps {len(plan['steps'])}")
    for s in plan["steps"]:
        if s["t"] == "launch":
            L.append(f"step launch {s['kernel']}")
        else:
            L.append(f"step levels {0 if s['order'] == 'forward' else 1} {len(s['sections'])}")
            for sec in s["sections"]:
                L.append(
                    f"section {b(sec['interval'][0])} {b(sec['interval'][1])} {len(sec['kernels'])} "
                    + " ".join(str(k) for k in sec["kernels"])
                )
    L.append("end")
    return "\n".join(L) + "\n"
