"""Streaming ("J-march") generator for PARALLEL computation blocks — the fast path of the backend.

Hand-written-style sm_100a kernel template, instantiated per `computation(PARALLEL)` block:

* one thread owns V consecutive I points (16 bytes: 4 x fp32 or 2 x fp64) -> every global access
  is a coalesced, vectorised LDG.128/STG.128 of an I-contiguous row segment; a warp owns 32*V
  consecutive I points of one row
* a warp *marches along J* over TJ output rows of one K level; every multi-stage temporary
  (`lap`, `flx`, `fly` in horizontal diffusion) and every input field lives in a small per-thread
  *register window* of rows (software pipeline: stage s works `lag_s` rows behind the newest
  loaded row), so each input row is read from memory once per tile and J-neighbours are register
  reuse, never a re-load
* I-neighbours come from the adjacent lanes with warp shuffles; the warp's first/last lanes are
  halo lanes (redundant compute instead of shared memory + barriers -> no __syncthreads at all)
* stages are fused: nothing but the API outputs (and temporaries another kernel needs) is stored
* masks / ternaries are predicated (straight-line code), exactly like the oracle's `np.where`

No tensor cores: the path is memory-bound point-wise arithmetic (12-16 B per cell, tens of flops).
The generic point generator (`codegen.py`) remains the fallback for everything outside this
template (K-sequential loops, while loops, variable K offsets, data dimensions, lower-dim fields).

Semantics: SURVEY.md §9 (numpy backend).  Reference component replaced: the GridTools multistage
with ij-caches, gtc/gtcpp/gtcpp_codegen.py:227-247.
"""

from __future__ import annotations

import math
from typing import Any, Dict, List, Optional, Tuple

from . import codegen as cg, ir as b2ir

CT = b2ir.CTYPE


class NotStreamable(Exception):
    pass


# ---------------------------------------------------------------------------------------------------
# analysis
# ---------------------------------------------------------------------------------------------------
class Value:
    """A row-streamed value: an input field at a K offset, or one version of a field written here."""

    def __init__(self, kind: str, field: str, dtype: str, *, dk: int = 0, ver: int = 0, stage: int = -1):
        self.kind, self.field, self.dtype, self.dk, self.ver, self.stage = kind, field, dtype, dk, ver, stage
        self.reads: List[Tuple[int, int, int]] = []  # (consumer stage, di, dj)
        self.store = False
        self.lag = 0
        self.window = 1
        self.nj = [0, 0]  # needed rows relative to the owned rows
        self.ni = [0, 0]  # needed columns relative to the owned columns

    @property
    def cname(self) -> str:
        base = cg._cname(self.field)
        return f"in_{base}_k{self.dk + 8}" if self.kind == "in" else f"v_{base}_{self.ver}"


class StreamKernel:
    def __init__(self, gen: "cg.Generator", interval, hes: List[dict], global_fields: set, opts: Dict[str, Any]):
        self.gen = gen
        self.ft = gen.ft
        self.interval = interval
        self.hes = hes
        self.global_fields = global_fields  # temporaries that other kernels touch
        (b0, o0), (b1, o1) = interval
        self.thin = b0 == b1 and 0 < o1 - o0 <= 2  # a section of one or two levels: few, short tasks; no shared-memory ring
        self.opts = {k: v for k, v in opts.items() if k not in ("tma", "min_blocks")} if self.thin else opts
        self.values: List[Value] = []
        self.vmap: Dict[Tuple, Value] = {}
        self.cur: Dict[str, Value] = {}
        self.stage_of_stmt: List = []
        self.direct: set = set()  # K-only fields read with uniform loads
        self.nstages = len(hes)

    # ---- eligibility + value discovery (pass 1) --------------------------------------------------
    def _decl(self, name: str) -> dict:
        return self.ft.entries[self.ft.index[name]]

    def _check_field(self, name: str, *, write: bool = False):
        d = self._decl(name)
        if d["data_dims"]:
            raise NotStreamable(f"field {name} has data dimensions")
        # a read-only IJ field streams like an IJK field whose K stride is 0; everything else must be IJK
        if not (all(d["dims"]) or (not write and d["dims"][0] and d["dims"][1])):
            raise NotStreamable(f"field {name} is not an IJK (or read-only IJ) field")
        if b2ir.ITEMSIZE[d["dtype"]] not in (4, 8):
            raise NotStreamable(f"field {name}: itemsize")

    def _in_value(self, name: str, dk: int) -> Value:
        key = ("in", name, dk)
        if key not in self.vmap:
            self._check_field(name)
            v = Value("in", name, self._decl(name)["dtype"], dk=dk)
            self.vmap[key] = v
            self.values.append(v)
        return self.vmap[key]

    def _new_version(self, name: str, stage: int) -> Value:
        self._check_field(name, write=True)
        n = sum(1 for v in self.values if v.kind == "tmp" and v.field == name)
        v = Value("tmp", name, self._decl(name)["dtype"], ver=n, stage=stage)
        self.vmap[("tmp", name, n)] = v
        self.values.append(v)
        return v

    def _resolve_read(self, node, stage: int) -> Tuple[Optional[Value], int, int]:
        off = node["off"]
        if isinstance(off, dict) or node.get("data_index"):
            raise NotStreamable("variable/absolute K offset or data index")
        di, dj, dk = off
        name = node["name"]
        d0 = self._decl(name)
        if not d0["dims"][0] and not d0["dims"][1] and not d0["data_dims"] and d0["kind"] == "api":
            # K-only (or 0-d) field: one value per level, the same for the whole warp -> plain uniform
            # load in the expression (read-only path), not a row stream
            self.direct.add(name)
            return "direct", 0, 0
        if name in self.cur:
            if dk != 0:
                raise NotStreamable("K-offset read of a field written in the same block")
            v = self.cur[name]
            if v.stage == stage and (di, dj) != (0, 0):
                raise NotStreamable("offset read of a value written in the same horizontal execution")
            return v, di, dj
        d = self._decl(name)
        if d["kind"] == "temp" and name not in self.global_fields:
            # kernel-local temporary read before any write: undefined in the oracle (np.empty)
            return None, di, dj
        return self._in_value(name, dk), di, dj

    def analyse(self) -> None:
        self.binding: Dict[int, Tuple[Optional[Value], int, int]] = {}  # id(expr node) -> read
        self.assign_ver: Dict[int, Tuple[Value, Optional[Value]]] = {}  # id(stmt) -> (new, prev)

        def visit_expr(e, stage):
            def fn(n):
                t = n["t"]
                if t == "field":
                    v, di, dj = self._resolve_read(n, stage)
                    self.binding[id(n)] = (v, di, dj)
                    if v is not None and v != "direct":
                        v.reads.append((stage, di, dj))
                elif t == "iter" and n["axis"] not in ("I", "J", "K"):
                    raise NotStreamable("iterator")

            b2ir.walk_exprs(e, fn)

        def visit_stmts(stmts, stage, masked):
            for s in stmts:
                t = s["t"]
                if t == "assign":
                    visit_expr(s["right"], stage)
                    left = s["left"]
                    if left["t"] == "field":
                        if isinstance(left["off"], dict) or left.get("data_index") or tuple(left["off"]) != (0, 0, 0):
                            raise NotStreamable("write with offset / data index")
                        name = left["name"]
                        prev = None
                        if masked:
                            if name in self.cur:
                                prev = self.cur[name]
                                prev.reads.append((stage, 0, 0))
                            else:
                                d = self._decl(name)
                                if d["kind"] == "api" or name in self.global_fields:
                                    prev = self._in_value(name, 0)
                                    prev.reads.append((stage, 0, 0))
                                    prev.pass_reads = getattr(prev, "pass_reads", 0) + 1  # only the "else" value of the masked write
                        new = self._new_version(name, stage)
                        self.assign_ver[id(s)] = (new, prev)
                        self.cur[name] = new
                elif t == "mask":
                    visit_expr(s["mask"], stage)
                    visit_stmts(s["body"], stage, True)
                elif t == "hregion":
                    visit_stmts(s["body"], stage, True)
                else:
                    raise NotStreamable(f"statement {t}")

        for si, he in enumerate(self.hes):
            visit_stmts(he["body"], si, False)
        # final versions of API fields / global temporaries are stored
        for name, v in self.cur.items():
            d = self._decl(name)
            if d["kind"] == "api" or name in self.global_fields:
                v.store = True
        if not any(v.store for v in self.values):
            raise NotStreamable("block stores nothing")
        # drop dead values (never read, never stored)? keep: harmless
        self.written = set(self.cur.keys())
        self._schedule()
        # in-place update through the stream: a field that this kernel stores may only be READ (in its incoming state)
        # at the very cells the reading thread owns.  A read at a K offset, or on halo lanes / halo rows (the value
        # is consumed at an IJ offset, directly or through an intermediate stage computed on an extended extent),
        # touches cells another warp or tile owns and may already have overwritten: a write-after-read race on the
        # device, wrong results on any multi-tile domain.  The point generator splits such loops at the hazard.
        # (Exempt: the incoming value is only the pass-through of ONE masked redefinition, `if m: f = x` — a cell a
        # neighbour has already stored holds `m ? x : old`, which is what this thread computes from either value.)
        for v in self.values:
            if v.kind == "in" and v.field in self.written and (v.dk != 0 or list(v.ni) != [0, 0] or list(v.nj) != [0, 0]):
                versions = sum(1 for x in self.values if x.kind == "tmp" and x.field == v.field)
                if v.dk == 0 and versions == 1 and len(v.reads) == getattr(v, "pass_reads", 0):
                    continue
                raise NotStreamable(f"field {v.field} is updated in place but read beyond the cells a thread owns")

    # ---- lags, windows, needed ranges (pass 2) -----------------------------------------------------
    def _schedule(self) -> None:
        n = self.nstages
        tmps = [v for v in self.values if v.kind == "tmp"]
        # ASAP lags per stage
        lag = [0] * n
        for s in range(n):
            for v in tmps:
                for (c, di, dj) in v.reads:
                    if c == s and v.stage != s:
                        lag[s] = max(lag[s], lag[v.stage] + dj)
        # ALAP: push producers as late as their consumers allow
        for s in reversed(range(n)):
            cons = [lag[c] - dj for v in tmps if v.stage == s for (c, di, dj) in v.reads if c != s]
            if cons:
                lag[s] = max(lag[s], min(cons))
        # re-validate (ALAP of a later stage cannot break earlier constraints, but be safe)
        for v in tmps:
            for (c, di, dj) in v.reads:
                if c != v.stage and lag[c] < lag[v.stage] + dj:
                    raise NotStreamable("lag constraints")
        base = min(lag)
        self.lag = [x - base for x in lag]
        # needed row / column ranges relative to the owned outputs, backwards
        need_j = [[None, None] for _ in range(n)]
        need_i = [[None, None] for _ in range(n)]

        def widen(r, lo, hi):
            r[0] = lo if r[0] is None else min(r[0], lo)
            r[1] = hi if r[1] is None else max(r[1], hi)

        for s in reversed(range(n)):
            for v in tmps:
                if v.stage != s:
                    continue
                if v.store:
                    widen(need_j[s], 0, 0)
                    widen(need_i[s], 0, 0)
                for (c, di, dj) in v.reads:
                    if c == s or need_j[c][0] is None:
                        continue
                    widen(need_j[s], need_j[c][0] + dj, need_j[c][1] + dj)
                    widen(need_i[s], need_i[c][0] + di, need_i[c][1] + di)
        for s in range(n):
            if need_j[s][0] is None:  # dead stage
                need_j[s], need_i[s] = [0, 0], [0, 0]
        self.need_j, self.need_i = need_j, need_i
        for v in self.values:
            if v.kind == "tmp":
                v.lag = self.lag[v.stage]
                v.nj, v.ni = list(need_j[v.stage]), list(need_i[v.stage])
                older = [self.lag[c] - dj for (c, di, dj) in v.reads if c != v.stage]
                v.window = max(1, (max(older) - v.lag + 1) if older else 1)
            else:
                if not v.reads:
                    continue
                v.lag = min(self.lag[c] - dj for (c, di, dj) in v.reads)
                v.window = max(self.lag[c] - dj for (c, di, dj) in v.reads) - v.lag + 1
                v.nj = [min(need_j[c][0] + dj for (c, di, dj) in v.reads), max(need_j[c][1] + dj for (c, di, dj) in v.reads)]
                v.ni = [min(need_i[c][0] + di for (c, di, dj) in v.reads), max(need_i[c][1] + di for (c, di, dj) in v.reads)]
        self.values = [v for v in self.values if v.kind == "tmp" or v.reads]
        # vector width: 16 bytes of the narrowest streamed field
        sizes = [b2ir.ITEMSIZE[v.dtype] for v in self.values if v.kind == "in" or v.store]
        if len(set(sizes)) != 1:
            raise NotStreamable("mixed item sizes among streamed API fields")
        self.V = 4 if min(sizes) == 4 else 2
        # register budget: a thread keeps (sum of window rows) x V values live; beyond ~40 32-bit
        # registers of window state the narrower vector (more resident warps) wins (measured on
        # horizontal diffusion: 78.8 % vs 66.6 % of HBM peak, profiles/README.md)
        pf = self._pf()
        rows = sum((v.window + (pf if v.kind == "in" else 0)) * (b2ir.ITEMSIZE[v.dtype] // 4) for v in self.values)
        vmax = self.V
        if self.V == 4 and rows * 4 > 40:
            self.V = 2
        if self.opts.get("vector_width"):
            self.V = min(vmax, int(self.opts["vector_width"]))
        V = self.V
        for v in self.values:
            for (c, di, dj) in v.reads:
                if abs(di) > 8:
                    raise NotStreamable("I offset too large")
        # halo lanes: every column at which ANY value is needed must be held by some lane of the warp — inputs read at
        # an I offset, but also intermediate stages that a later stage reads at an I offset (a stage read at +1 needs a
        # compute-only lane to the right even when no input is read at a positive offset; found by tools/fuzz_codegen.py)
        ranged = [v for v in self.values if getattr(v, "ni", None) is not None]
        lreach = max([0] + [-v.ni[0] for v in ranged])
        rreach = max([0] + [v.ni[1] for v in ranged])
        self.hl, self.hr = math.ceil(lreach / V), math.ceil(rreach / V)
        self.SQ = 32 - self.hl - self.hr
        # bulk-async variant: a warp row is copied by ONE cp.async.bulk (16-byte aligned source, size a multiple of
        # 16): when a vector is only 8 bytes (fp32 x 2) the segments must start at even vector indices
        isz = min(sizes)
        self.tma_tensor = self.opts.get("tma_mode", "tensor") == "tensor"
        # (both copy flavours: a tensor-map copy whose box starts at an 8-byte-aligned element is an illegal
        #  instruction on the device — measured, profiles/README.md r02e)
        self.qalign = 2 if (self.opts.get("tma") and V * isz == 8) else 1
        if self.qalign == 2 and self.SQ % 2:
            self.SQ -= 1
        self._choose_windows()
        self.TJ = int(self.opts.get("tile_j", 64))
        # a section of one or two levels (the top / bottom level of a K-dependent computation) has few tasks: shorter J
        # tiles, or its launch is a fraction of a wave of warps that each march 64 rows (75 us for one 4096 x 512 plane)
        if self.thin:
            self.TJ = int(self.opts.get("thin_tile_j", 8))
        self.NW = int(self.opts.get("warps", 4))

    def _pf(self) -> int:
        """Extra window rows of the input streams (loads issued that many march steps before first use).  The
        bulk-async variant (`tma`) needs none: its look-ahead lives in the shared-memory ring, not in registers."""
        if "prefetch" in self.opts:
            return int(self.opts["prefetch"])
        # one extra row per input stream pays while the streams are few (hdiff 2, upwind5 3: 0.73 -> 0.83 of the HBM peak in
        # round 1); a kernel that streams many inputs (pressure gradient: 9 row streams, divergence: 6) is short of
        # registers instead — 90 registers, 1.18 ms with the extra rows, 1.03 ms without (profiles/README.md r02o)
        n_in = sum(1 for v in self.values if v.kind == "in")
        return 0 if (self.opts.get("tma") or n_in > 4) else 1

    def _choose_windows(self) -> None:
        """Pick the rotation period U and the final register-window sizes.

        Inputs get `prefetch` extra rows (loads issued that many march steps before first use ->
        more bytes in flight per warp).  A window that divides U rotates by register renaming in the
        U-times unrolled march loop (no moves); the others shift (W-1 moves per element and step)."""
        pf = self._pf()
        req = {}
        for v in self.values:
            req[id(v)] = v.window + (pf if v.kind == "in" else 0)
        best = None
        # the march loop is unrolled P times: its code must stay well inside the 32 KB instruction cache of an SM (L1.5),
        # warps of a CTA sit at different places of the loop.  Measured (profiles/README.md r02m): upwind5 with P = 8 is a
        # 30 KB loop — 49 % instruction-cache hit rate, "no instruction" the top stall, 2.8 ms; P = 4 with phi's window
        # shifting: 1.7 ms.  Rough size of one step: 16 bytes x V x (operations + shuffles/loads/stores).
        ops = [0]

        def count(n):
            if n["t"] in ("binary", "unary", "ternary", "call", "cast"):
                ops[0] += 6 if (n["t"] == "binary" and n["op"] == "/") else (12 if n["t"] == "call" else 1)
            elif n["t"] == "field":
                ops[0] += 1

        for he in self.hes:
            b2ir.walk_exprs(he["body"], count)
        step_bytes = 16 * (self.V * ops[0] + 8 * len([v for v in self.values if v.kind == "in" or v.store]))
        forced = int(self.opts.get("period", 0) or 0)  # option `period`: force the unroll factor (device sweeps)
        for P in (1, 2, 3, 4, 6, 8):  # (8: a 7-row window + 1 prefetch row rotates instead of shifting: upwind5's phi)
            if forced and P != forced:
                continue
            divs = [d for d in range(1, P + 1) if P % d == 0]
            cost, assign = 0, {}
            for v in self.values:
                r = req[id(v)]
                fit = [d for d in divs if d >= r]
                if fit:
                    assign[id(v)] = (fit[0], False)
                    cost += fit[0]
                else:
                    assign[id(v)] = (r, True)
                    cost += 3 * r
            cost += 0.25 * P  # code size
            cost += max(0.0, (P * step_bytes - 14 * 1024) / 1024.0) * 2.0  # instruction-cache pressure beyond ~14 KB of loop
            if best is None or cost < best[0]:
                best = (cost, P, assign)
        _, self.U, assign = best
        for v in self.values:
            w, shift = assign[id(v)]
            if v.kind == "in":
                v.lag -= w - v.window  # loads run ahead of first use
            v.window, v.shift = w, shift

    def declared_extent(self, name: str) -> List[List[int]]:
        """Authoritative halo of a field = what validation / scratch allocation guarantee to exist:
        API field -> boundary from field_info; temporary -> its extent."""
        d = self._decl(name)
        if d["kind"] == "temp":
            return d["extent"]
        b = self.gen.st["field_info"][name]["boundary"]
        return [[-b[0][0], b[0][1]], [-b[1][0], b[1][1]]]

    # ---- emission (pass 3) ----------------------------------------------------------------------------
    def reg(self, v: Value, slot: int, e: int) -> str:
        return f"{v.cname}_s{slot}_{e}"

    def _slot_of(self, v: Value, consumer_lag: int, dj: int, phase: int) -> int:
        m = consumer_lag - dj - v.lag  # rows behind the newest row of v
        assert 0 <= m < v.window, (v.cname, m, v.window)
        return m if v.shift else (phase - m) % v.window

    def _new_slot(self, v: Value, phase: int) -> int:
        return 0 if v.shift else phase % v.window

    def emit(self) -> int:
        gen, V, U = self.gen, self.V, self.U
        _HINTS["ld"] = "__ldcs" if self.opts.get("ldcs", False) else "__ldg"
        _HINTS["st"] = "__stcs" if self.opts.get("stcs", False) else ""
        name = gen._kname("stream")
        A = "A"
        L: List[str] = []
        w = L.append
        nthreads = 32 * self.NW
        minb = int(self.opts.get("min_blocks", 0))
        lb = f"__launch_bounds__({nthreads}, {minb})" if minb else f"__launch_bounds__({nthreads})"
        w(f'extern "C" __global__ void {lb} {name}(const __grid_constant__ Args A) {{')
        w(f"  constexpr int V = {V}, SQ = {self.SQ}, HL = {self.hl}, TJ = {self.TJ}, NW = {self.NW};")
        w("  const int lane = threadIdx.x;")
        # divisions by launch-invariant divisors (scalar parameters / literals: grid spacings, time steps): the divisor and
        # its correctly rounded reciprocal are set up once per thread HERE (the lines are inserted when the march steps
        # have been emitted), every cell then pays 3 FP instructions instead of an IEEE division sequence (b200::DivInv)
        self.div_hoisted: Dict[str, str] = {}
        div_marker = len(L)
        # owned output box of this launch: union of the storing stages' extents
        st_stages = sorted({v.stage for v in self.values if v.store})
        ei0 = min(self.hes[s]["extent"][0][0] for s in st_stages)
        ei1 = max(self.hes[s]["extent"][0][1] for s in st_stages)
        ej0 = min(self.hes[s]["extent"][1][0] for s in st_stages)
        ej1 = max(self.hes[s]["extent"][1][1] for s in st_stages)
        self.store_ext = [[ei0, ei1], [ej0, ej1]]
        w(f"  const int X0 = {A}.g.i_lo + ({ei0}), X1 = {A}.g.i_hi + ({ei1});")
        w(f"  const int Y0 = {A}.g.j_lo + ({ej0}), Y1 = {A}.g.j_hi + ({ej1});")
        w("  const int QXF = (X0 >= 0) ? X0 / V : -((-X0 + V - 1) / V);")
        if self.qalign == 2:
            w("  const int QX0 = QXF - (((QXF - HL) % 2 + 2) % 2);   // lane 0 of every segment sits on a 16-byte boundary")
        else:
            w("  const int QX0 = QXF;")
        w("  const int nseg = ((X1 - QX0 * V) + SQ * V - 1) / (SQ * V);")
        w("  const int ntj = (Y1 - Y0 + TJ - 1) / TJ;")
        if self.opts.get("tma") or self.opts.get("uniform_task", True):
            # warp-uniform by construction (a shuffle from lane 0): everything derived from it — task, segment, tile, level,
            # march step, row base addresses, ring slot — stays in uniform registers: hdiff 64 -> 44 registers (32 -> 44 resident
            # warps per SM) and 0.85 -> 0.91 of the HBM peak (profiles/README.md r02k); the bulk-copy operands of the `tma`
            # variant need no per-lane waterfall loop.  `uniform_task=False` keeps the plain threadIdx.y form.
            w("  const int wy = __shfl_sync(0xffffffffu, (int)threadIdx.y, 0);")
        else:
            w("  const int wy = threadIdx.y;")
        w("  const long long task = (long long)blockIdx.x * NW + wy;")
        w(f"  const int nk = {A}.g.k_hi - {A}.g.k_lo;")
        w("  if (task >= (long long)nseg * ntj * nk) return;")
        halo_wait = bool(self.opts.get("halo_wait", False))

        # task order.  Default: segment fastest, level slowest (plane by plane).  A kernel that reads a field at K offsets
        # (pp[0,0,+-1] of the pressure gradient) touches every plane of that field from three levels: with the LEVEL
        # fastest, the warps of a CTA and of its neighbours work on consecutive levels of the same rows at the same time,
        # so the second and third read of a row hit L1 / L2 instead of HBM (plane by plane the re-use distance is two
        # whole planes of every streamed field: 37 instead of 28 B/cell of DRAM traffic, profiles/README.md r02m)
        kfast = self.opts.get("k_order", "auto")
        kfast = any(v.kind == "in" and v.dk != 0 for v in self.values) if kfast == "auto" else bool(kfast)
        self.kfast = kfast
        # (development switches of the halo_wait form: task order and wait block separately)
        halo_order = halo_wait and bool(self.opts.get("halo_order", True))
        halo_block = halo_wait and bool(self.opts.get("halo_block", True))
        if halo_order and kfast:
            w("  int seg, tj, kk;")
            w("  if (ntj >= 3) {")
            w("    const long long n_int = (long long)nseg * (ntj - 2) * nk;   // interior tiles first")
            w("    if (task < n_int) { kk = (int)(task % nk); const long long r = task / nk; seg = (int)(r % nseg); tj = 1 + (int)(r / nseg); }")
            w("    else { const long long tb = task - n_int; kk = (int)(tb % nk); const long long r = tb / nk; seg = (int)(r % nseg); tj = ((r / nseg) & 1) ? ntj - 1 : 0; }")
            w("  } else { kk = (int)(task % nk); const long long r = task / nk; seg = (int)(r % nseg); tj = (int)(r / nseg); }")
            w(f"  const int k = {A}.g.k_lo + kk;")
        elif kfast:
            w(f"  const int k = {A}.g.k_lo + (int)(task % nk);")
            w("  const int seg = (int)((task / nk) % nseg);")
            w("  const int tj = (int)(task / ((long long)nk * nseg));")
        elif halo_order and self.opts.get("halo_lean", False):
            # the same order — interior tiles first, the two boundary tiles of every level last — decoded with selects instead
            # of branches: the task coordinates stay visibly warp-uniform for the compiler (no 10-register penalty)
            w("  const bool reorder = ntj >= 3;")
            w("  const long long n_int = reorder ? (long long)nseg * (ntj - 2) * nk : 0LL;   // interior tiles first")
            w("  const bool bnd = reorder && task >= n_int;")
            w("  const long long tsk = bnd ? task - n_int : task;")
            w("  const int seg = (int)(tsk % nseg);")
            w("  const long long q = tsk / nseg;")
            w("  const int rows = bnd ? 2 : (reorder ? ntj - 2 : ntj);   // tiles per level in this part of the order")
            w("  const int qq = (int)(q % rows);")
            w("  const int tj = bnd ? (qq ? ntj - 1 : 0) : (reorder ? 1 : 0) + qq;")
            w(f"  const int k = {A}.g.k_lo + (int)(q / rows);")
        elif halo_order:
            # multi-GPU J slabs with the peer-memory halo exchange: the first and the last J tile read rows the neighbours
            # push into this rank's halo while this kernel is already running -> those tiles come LAST in the task order
            # (all levels), behind every interior tile, so that by the time they start their halo has normally arrived
            w("  int seg, tj, kk;")
            w("  if (ntj >= 3) {")
            w("    const long long n_int = (long long)nseg * (ntj - 2) * nk;   // interior tiles first")
            w("    if (task < n_int) { seg = (int)(task % nseg); tj = 1 + (int)((task / nseg) % (ntj - 2)); kk = (int)(task / ((long long)nseg * (ntj - 2))); }")
            w("    else { const long long tb = task - n_int; seg = (int)(tb % nseg); tj = ((tb / nseg) & 1) ? ntj - 1 : 0; kk = (int)(tb / (2LL * nseg)); }")
            w("  } else { seg = (int)(task % nseg); tj = (int)((task / nseg) % ntj); kk = (int)(task / ((long long)nseg * ntj)); }")
            w(f"  const int k = {A}.g.k_lo + kk;")
        else:
            w("  const int seg = (int)(task % nseg);")
            w("  const int tj = (int)((task / nseg) % ntj);")
            w(f"  const int k = {A}.g.k_lo + (int)(task / ((long long)nseg * ntj));")
        k0 = cg._bound(self.interval[0], f"{A}.g.nK")
        k1 = cg._bound(self.interval[1], f"{A}.g.nK")
        w(f"  if (k < {k0} || k >= {k1}) return;")
        w("  const int g0w = (QX0 + seg * SQ - HL) * V;   // first I index of the warp's row segment (warp-uniform)")
        w("  const int g0 = g0w + lane * V;   // first I index of this lane's vector")
        w("  const bool own = (lane >= HL) && (lane < HL + SQ);")
        w("  const int T_lo = Y0 + tj * TJ;")
        w("  const int T_hi = (T_lo + TJ < Y1) ? T_lo + TJ : Y1;")
        if halo_block:
            ins_nj = [v.nj for v in self.values if v.kind == "in"]
            lo_rows, hi_rows = min([0] + [n[0] for n in ins_nj]), max([0] + [n[1] for n in ins_nj])
            w(f"  if ({A}.g.halo_epoch) {{   // tiles that read halo rows wait until the neighbour's rows of this step have landed")
            w(f"    const bool need_lo = {A}.g.halo_flag_lo && (T_lo + ({lo_rows}) < 0);")
            w(f"    const bool need_hi = {A}.g.halo_flag_hi && (T_hi + ({hi_rows}) > {A}.g.nJ);")
            w("    if (need_lo || need_hi) {")
            # (`halo_lean`: task decode by selects + a wait without time limit: 40 instead of 53 registers for hdiff, the
            #  plain kernel has 44 — static evidence only, not yet measured on a device: opt-in)
            wf = "wait_flag_nolimit" if self.opts.get("halo_lean", False) else "wait_flag"
            w("      if (lane == 0) {")
            w(f"        if (need_lo) b200::{wf}({A}.g.halo_flag_lo, {A}.g.halo_epoch);")
            w(f"        if (need_hi) b200::{wf}({A}.g.halo_flag_hi, {A}.g.halo_epoch);")
            w("      }")
            w("      __syncwarp();")
            if self.opts.get("tma"):
                # the neighbour's rows were written through the generic proxy (NVLink stores) and acquired above by lane 0; the
                # bulk copies of this warp read them through the ASYNC proxy: order the two proxies before the first copy
                w("      b200::fence_proxy_async_global();")
            w("    }")
            w("  }")
        # window registers start at ONE, not zero: rows and lanes outside the tile are never loaded, whatever is computed
        # from them is never stored, but a zero dividend would send the warp's quotient group through the IEEE fallback
        # (upwind5: 17 % of all executed instructions were fallback divisions of prologue rows, profiles/README.md r02n)
        for v in self.values:
            ct = CT[v.dtype]
            for s in range(v.window):
                w("  " + f"{ct} " + ", ".join(f"{self.reg(v, s, e)} = ({ct})1" for e in range(V)) + ";")
        self.row_guard: Dict[str, bool] = {}
        for v in self.values:
            if v.kind == "in" or v.store:
                n = self.ft.index[v.field]
                ct = CT[v.dtype]
                w(f"  const {ct}* const p_{v.cname} = (const {ct}*){A}.f[{n}].p + (long long)(k + ({v.dk})) * {A}.f[{n}].s[2];"
                  if v.kind == "in" else
                  f"  {ct}* const p_{v.cname} = ({ct}*){A}.f[{n}].p + (long long)k * {A}.f[{n}].s[2];")
                w(f"  const bool vec_{v.cname} = {A}.f[{n}].vec != 0;")
                w(f"  const long long sj_{v.cname} = {A}.f[{n}].s[1], si_{v.cname} = {A}.f[{n}].s[0];")
                if v.kind == "in":
                    de = self.declared_extent(v.field)
                    lo = f"{A}.g.i_lo + ({max(ei0 + v.ni[0], de[0][0])})"
                    hi = f"{A}.g.i_hi + ({min(ei1 + v.ni[1], de[0][1])})"
                    w(f"  const int rlo_{v.cname} = {A}.g.j_lo + ({de[1][0]}), rhi_{v.cname} = {A}.g.j_hi + ({de[1][1]});")
                    # rows the tile needs may exceed what is declared valid only if the need analysis is
                    # coarser than the extent analysis; then the steady loop keeps the row guard
                    self.row_guard[v.cname] = not (ej0 + v.nj[0] >= de[1][0] and ej1 + v.nj[1] <= de[1][1])
                else:
                    e = self.hes[v.stage]["extent"]
                    lo = f"{A}.g.i_lo + ({e[0][0]})"
                    hi = f"{A}.g.i_hi + ({e[0][1]})"
                    self.row_guard[v.cname] = not (e[1][0] == ej0 and e[1][1] == ej1)
                w(f"  const int clo_{v.cname} = {lo}, chi_{v.cname} = {hi};")
                w(f"  const bool full_{v.cname} = vec_{v.cname} && g0 >= clo_{v.cname} && g0 + V <= chi_{v.cname};")
                w(f"  const bool any_{v.cname} = g0 + V > clo_{v.cname} && g0 < chi_{v.cname};")
                # branch-free handling of partially valid vectors in the steady loop (unit I stride)
                gate = "" if v.kind == "in" else "own && "
                for e in range(V):
                    w(f"  const bool pe_{v.cname}_{e} = {gate}vec_{v.cname} && !full_{v.cname} && g0 + {e} >= clo_{v.cname} && g0 + {e} < chi_{v.cname};")
                if v.kind != "in":
                    w(f"  const bool sfull_{v.cname} = own && full_{v.cname};")
                    w(f"  const bool wpart_{v.cname} = __any_sync(0xffffffffu, " + " || ".join(f"pe_{v.cname}_{e}" for e in range(V)) + ");")
                else:
                    # warp-uniform: does any lane of this warp hold a partially valid vector of this stream?
                    w(f"  const bool wpart_{v.cname} = __any_sync(0xffffffffu, " + " || ".join(f"pe_{v.cname}_{e}" for e in range(V)) + ");")
        w("  const bool allvec = " + " && ".join(f"vec_{v.cname}" for v in self.values if v.kind == "in" or v.store) + ";")
        # warp-uniform: no lane of this warp holds a partially valid vector of any stream -> the
        # steady loop needs one predicate per stream and nothing else
        w("  const bool wpure = !(" + " || ".join(f"wpart_{v.cname}" for v in self.values if v.kind == "in" or v.store) + ");")
        pfd = int(self.opts.get("l2_prefetch", 2))
        if pfd > 0:
            for v in self.values:
                if v.kind == "in":
                    w(f"  const bool pfl_{v.cname} = (lane & 3) == 0 && full_{v.cname};   // lanes that issue L2 prefetches")
                    w(f"  const int pfe_{v.cname} = T_hi + ({v.nj[1] + v.lag - pfd});      // last step (exclusive) that may prefetch")
        # march ranges (in step space, relative to the tile rows)
        first_terms = [self.need_j[s][0] + self.lag[s] for s in range(self.nstages)] + [
            v.nj[0] + v.lag for v in self.values if v.kind == "in"
        ]
        last_terms = [self.need_j[s][1] + self.lag[s] for s in range(self.nstages)]
        t_first, t_last = min(first_terms), max(last_terms)
        # steady state: every load row is inside the tile's needed rows, every store row is owned
        lo_terms = [v.nj[0] + v.lag for v in self.values if v.kind == "in"] + [v.lag for v in self.values if v.store]
        hi_terms = [v.nj[1] + v.lag for v in self.values if v.kind == "in"] + [v.lag for v in self.values if v.store]
        a, b = max(lo_terms), min(hi_terms)
        n_pro = a - t_first
        w(f"  int t = T_lo + ({t_first});")
        w(f"  const int t_end = T_hi + ({t_last});   // exclusive")
        streams = [v for v in self.values if v.kind == "in" or v.store]
        ph0 = n_pro % U
        # static pitch: the row pitch (J stride, in elements) of every streamed field is a compile-time
        # constant -> row addresses inside a trip are immediates off ONE running pointer per stream
        # (no per-step 64-bit address arithmetic); checked at run time, other pitches take the general loop
        self.SJ = int(self.opts.get("static_pitch", 0) or 0)
        pitch_ok = ""
        if self.SJ:
            w(f"  constexpr long long SJ = {self.SJ};")
            w("  const bool spitch = " + " && ".join(f"sj_{v.cname} == SJ" for v in streams) + ";")
            pitch_ok = " && spitch"

        self._lag_of = {v.cname: v.lag for v in streams}

        # static pitch: rows are addressed as p + (t*SJ + g0) + immediate (one shared offset, 2 registers
        # less per stream) unless `row_pointers` asks for one running pointer per stream (fewer instructions)
        self._useq = not self.SJ or bool(self.opts.get("row_pointers", False))

        def running_pointers():
            if not self._useq:
                return
            for v in streams:
                ct = CT[v.dtype]
                const = "const " if v.kind == "in" else ""
                w(f"  {const}{ct}* q_{v.cname} = p_{v.cname} + (long long)(t - ({v.lag})) * sj_{v.cname} + g0;")

        def steady_loop(mode, cond, bound):
            w(f"  if ({cond}) for (; t + {U} <= {bound}; t += {U}) {{   // steady loop ({mode})")
            w(f"    B200_TRACE({3 if mode == 'interior' else (0 if mode == 'pure' else 1)});")
            if not self._useq:
                w("    const long long toff = (long long)t * SJ + g0;   // this trip's row offset, shared by all streams")
            for u in range(U):
                w(f"    {{  // steady step, rotation phase {(ph0 + u) % U}")
                w(f"      const int tt = t + {u};")
                L.extend(self._emit_step((ph0 + u) % U, fast=mode, u=u))
                w("    }")
            if self.SJ and self._useq:
                for v in streams:
                    w(f"    q_{v.cname} += {U} * SJ;")
            w("  }")

        def row_at(v, ahead: int, trip_offset: str) -> str:
            """Address of lane's vector in the row of stream v that march step `t + ahead` loads (outside the trip body)."""
            c = v.cname
            if self.SJ and not self._useq:
                return f"(p_{c} + {trip_offset} + ({ahead - v.lag}) * SJ)"
            pitch = "SJ" if self.SJ else f"sj_{c}"
            return f"(q_{c} + {ahead} * {pitch})" if ahead else f"q_{c}"

        def steady_loop_tma(cond):
            """Bulk-async steady loop: the ring has D slots, a slot holds the input rows of ONE TRIP (U march steps x NS
            streams); one mbarrier, one wait, one warp sync and one refill (by lane 0) per trip."""
            D, NS, RB, UT = self.tma, len(ins), self.tma_rb, self.tma_ut
            slot_bytes = NS * UT * RB

            def issue(slot_expr: str, bar: str, first_step: str, pad: str):
                w(f"{pad}b200::mbar_expect_tx({bar}, {slot_bytes});")
                if self.tma_tensor:
                    # ONE tensor-map copy per stream and trip: a box of 32*V elements x UT rows of level k (+dk), rows
                    # packed back to back in the slot (coordinates are array indices: domain index + origin)
                    for n, v in enumerate(ins):
                        m = gen.tmap_index(v.field, 32 * V, UT)
                        w(f"{pad}b200::tma_load_3d({slot_expr} + {n * UT * RB}, &A.tm[{m}], g0w + A.tmo[{m}][0], "
                          f"({first_step} - ({v.lag})) + A.tmo[{m}][1], (k + ({v.dk}) + A.tmo[{m}][2]) * A.tmo[{m}][3], {bar});")
                    return
                for n, v in enumerate(ins):
                    pitch = "SJ" if self.SJ else f"sj_{v.cname}"
                    w(f"{pad}{{ const {CT[v.dtype]}* src = p_{v.cname} + (long long)({first_step} - ({v.lag})) * {pitch} + g0w;")
                    for u in range(UT):
                        w(f"{pad}  b200::bulk_g2s({slot_expr} + {(n * UT + u) * RB}, src + {u} * {pitch}, {RB}, {bar});")
                    w(f"{pad}}}")

            w(f"  if ({cond}) {{   // steady loop (interior, bulk-async ring: {D} slots of {UT} rows x {NS} streams per warp)")
            w(f"    const int t_tma_end = (t_int > t) ? t + ((t_int - t) / {UT}) * {UT} : t;   // first step after the steady trips")
            w("    if (t_tma_end > t) {")
            w(f"      unsigned char* rpw = &b200_ring[wy][0];   // the current slot of this warp's ring (warp-uniform)")
            w(f"      unsigned char* rp = rpw + lane * {RB // 32};   // this lane's vector in it")
            w("      int sl = 0;")
            w("      unsigned par = 0u;")
            w("      const bool issuer = b200::elect_one();   // one lane issues the copies")
            w("      if (issuer) {")
            w(f"        for (int d = 0; d < {D}; ++d) b200::mbar_init(&b200_bar[wy][d], 1);")
            w("        b200::mbar_fence_init();")
            for d in range(D):
                w(f"        if (t + {d * UT} < t_tma_end) {{")
                issue(f"rpw + {d * slot_bytes}", f"&b200_bar[wy][{d}]", f"t + {d * UT}", "          ")
                w("        }")
            w("      }")
            w("      __syncwarp();")
            w(f"      for (; t + {UT} <= t_int; t += {UT}) {{")
            w("        B200_TRACE(3);")
            w("        B200_TRACE(5);")
            if not self._useq:
                w("        const long long toff = (long long)t * SJ + g0;   // this trip's row offset, shared by all streams")
            w("        b200::mbar_wait(&b200_bar[wy][sl], par);   // the rows of this trip have landed")
            for u in range(UT):
                w(f"        {{  // steady step, rotation phase {(ph0 + u) % U}")
                w(f"          const int tt = t + {u};")
                L.extend(self._emit_step((ph0 + u) % U, fast="interior", u=u))
                w("        }")
            w("        __syncwarp();   // every lane has read the slot: it may be refilled")
            w(f"        if (issuer && t + {D * UT} < t_tma_end) {{")
            if self.opts.get("tma_fence", True):
                w("          b200::fence_async_smem();")
            issue("rpw", "&b200_bar[wy][sl]", f"t + {D * UT}", "          ")
            w("        }")
            if self.SJ and self._useq:
                for v in streams:
                    w(f"        q_{v.cname} += {UT} * SJ;")
            w(f"        ++sl; rp += {slot_bytes}; rpw += {slot_bytes};")
            w(f"        if (sl == {D}) {{ sl = 0; rp -= {D * slot_bytes}; rpw -= {D * slot_bytes}; par ^= 1u; }}")
            w("      }")
            w("    }")
            w("  }")

        def prologue(fast):
            # steps until the steady state starts (compile-time count)
            for n in range(n_pro):
                w(f"  if (t < t_end) {{  // prologue step {n}")
                w("      const int tt = t;")
                L.extend(self._emit_step(n % U, fast=fast))
                w("  }")
                w("  ++t;")

        def tail(fast):
            # remaining steady rows + epilogue, same phase sequence
            w(f"  for (; t < t_end; t += {U}) {{")
            for u in range(U):
                w(f"    if (t + {u} < t_end) {{  // tail step, rotation phase {(ph0 + u) % U}")
                w(f"      B200_TRACE({4 if fast else 2});")
                w(f"      const int tt = t + {u};")
                L.extend(self._emit_step((ph0 + u) % U, fast=fast))
                w("    }")
            w("  }")

        # interior warps: all 32 lanes hold fully valid vectors of every stream (all but the last I
        # segment of a row) and every row the tile needs exists -> the WHOLE march runs without
        # per-lane predicates: prologue / epilogue steps only test warp-uniform row ranges, steady
        # trips have unconditional loads and run while every L2 prefetch is allowed
        interior = self.opts.get("interior_loop", False) and self.opts.get("pure_loop", True) and not any(self.row_guard.values())
        # interior_loop="steady": only the steady trips of interior warps are specialised (fewer registers)
        steady_only = interior and self.opts.get("interior_loop") == "steady"
        # bulk-async variant (option tma=D): the input rows of interior warps come through a per-warp ring of D rows in
        # shared memory, filled by cp.async.bulk copies that one lane issues D march steps ahead (mbarrier per slot);
        # the lanes read their vectors with LDS.  Look-ahead costs shared memory instead of registers and L2 prefetches.
        self.tma = 0
        ins = [v for v in self.values if v.kind == "in"]
        if int(self.opts.get("tma", 0) or 0) and interior and ins:
            isz = b2ir.ITEMSIZE[ins[0].dtype]
            self.tma_rb = 32 * V * isz  # bytes of one warp row
            # ring depth in TRIPS; a trip = the smallest multiple of the rotation period with at least `tma_rows` steps.
            # The ring must fit the shared-memory budget of a CTA (static allocation, and small enough that shared
            # memory does not cap the resident CTAs below what the registers allow): shorter trips first, then fewer slots.
            budget = int(self.opts.get("tma_smem_kb", 40)) * 1024
            want_d = max(2, int(self.opts["tma"]))
            want_m = max(1, -(-int(self.opts.get("tma_rows", 6)) // U))
            fits = [(d, m) for d in range(want_d, 1, -1) for m in range(want_m, 0, -1)
                    if self.NW * d * (len(ins) * m * U * self.tma_rb + 8) <= min(budget, 48 * 1024)]
            if fits:
                D, m = fits[0]
                self.tma, self.tma_ut = D, m * U
                slot = len(ins) * self.tma_ut * self.tma_rb
                w(f"  __shared__ __align__(128) unsigned char b200_ring[NW][{D * slot}];   // [warp][slot][stream][step of the trip][row bytes]")
                w(f"  __shared__ __align__(8) unsigned long long b200_bar[NW][{D}];")
        if interior:
            w("  const bool winterior = __all_sync(0xffffffffu, " + " && ".join(
                f"full_{v.cname}" if v.kind == "in" else f"(!own || full_{v.cname})" for v in streams) + ");")
            lim = [f"T_hi + ({b})"] + ([f"pfe_{v.cname}" for v in self.values if v.kind == "in"] if pfd > 0 and not self.tma else [])
            w("  int t_int = " + lim[0] + ";")
            for x in lim[1:]:
                w(f"  t_int = t_int < {x} ? t_int : {x};")
            if pfd > 0:
                w("  const bool pf_lane = (lane & 3) == 0;")
        if interior and not steady_only:
            w(f"  if (allvec && wpure && winterior{pitch_ok}) {{")
            prologue("igeneral")
            running_pointers()
            if self.tma:
                steady_loop_tma("true")
            else:
                steady_loop("interior", "true", "t_int")
            tail("igeneral")
            w("  } else {")
        prologue(False)
        running_pointers()
        modes = [("pure", "allvec && wpure" + pitch_ok, f"T_hi + ({b})")]
        if steady_only:
            modes.insert(0, ("interior", "allvec && wpure && winterior" + pitch_ok, "t_int"))
        if not self.opts.get("pure_loop", True):
            modes = [("fastall", "allvec" + pitch_ok, f"T_hi + ({b})")]  # one steady loop for all warps, uniform branch per load
        # a second steady loop for warps that do hold partially valid vectors costs ~18 registers
        # (72 -> 90 for horizontal diffusion); by default those few edge warps use the general loop
        if self.opts.get("edge_loop", False) and self.opts.get("pure_loop", True):
            modes.append(("fast", "allvec && !wpure" + pitch_ok, f"T_hi + ({b})"))
        for mode, cond, bound in modes:
            if mode == "interior" and self.tma:
                steady_loop_tma(cond)
            else:
                steady_loop(mode, cond, bound)
        tail(False)
        if interior and not steady_only:
            w("  }")
        w("}")
        L[div_marker:div_marker] = [f"  const auto {nm} = b200::div_inv_make({cx});" for cx, nm in self.div_hoisted.items()]
        gen.src.append("\n".join(L))
        gen.live |= {v.field for v in self.values if v.kind == "in" or v.store} | self.direct
        gen.kernels.append(
            {
                "name": name, "kind": "stream", "block": [32, self.NW, 1], "tile": [self.SQ * V, self.TJ, V],
                "extent": self.store_ext, "k_lo": self.interval[0], "k_hi": self.interval[1], "smem": 0,
                "qshift": self.qalign - 1, "tma": self.tma, "tma_mode": ("tensor" if self.tma_tensor else "bulk") if self.tma else None, "vector": V, "period": U, "windows": {v.cname: (v.window, "shift" if v.shift else "rot") for v in self.values},
            }
        )  # fmt: skip
        return len(gen.kernels) - 1

    # one march step at rotation phase `phase`
    def _emit_step(self, phase: int, fast, u: int = 0) -> List[str]:
        V = self.V
        L: List[str] = []
        ind = "      "
        self._u = u
        # 0. shifting windows (values whose window does not divide the rotation period)
        for v in self.values:
            if v.shift and v.window > 1:
                for s in range(v.window - 1, 0, -1):
                    L.append(ind + " ".join(f"{self.reg(v, s, e)} = {self.reg(v, s - 1, e)};" for e in range(V)))
        # 1. loads of the newest row of every input stream
        if fast == "interior" and getattr(self, "tma", 0):
            L.extend(self._emit_ring_loads(phase, u))
        for v in self.values:
            if v.kind != "in" or (fast == "interior" and getattr(self, "tma", 0)):
                continue
            slot = self._new_slot(v, phase)
            c = v.cname
            ct = CT[v.dtype]
            ro = v.field not in self.written
            regs = [self.reg(v, slot, e) for e in range(V)]
            ldf = "__ldg" if ro else "*"
            if fast and fast != "igeneral":
                guard = f"R >= rlo_{c} && R < rhi_{c}" if self.row_guard[c] else ""
                L.append(f"{ind}{{")
                if guard:
                    L.append(f"{ind}  const int R = tt - ({v.lag}); const bool rok = {guard};")
                g = "rok && " if guard else ""
                def edge_lines(pad, qa):
                    # some lane partially valid: loads go to fresh temporaries and are merged with
                    # selects, so no load waits on another one's destination registers
                    tmpv = [f"tv{e}" for e in range(V)]
                    tmps = [f"ts{e}" for e in range(V)]
                    out = [f"{pad}{ct} " + ", ".join(f"{t} = {regs[e]}" for e, t in enumerate(tmpv)) + ";"]
                    out.append(f"{pad}if ({g}full_{c}) {_vec_load(ct, V, tmpv, qa, ro)}")
                    for e in range(V):
                        out.append(f"{pad}const {ct} {tmps[e]} = ({g}pe_{c}_{e}) ? {ldf}({qa} + {e}) : {regs[e]};")
                    for e in range(V):
                        out.append(f"{pad}{regs[e]} = full_{c} ? {tmpv[e]} : {tmps[e]};")
                    return out

                qa = self._row_addr(c)
                if fast == "interior":
                    # every lane of the warp holds a fully valid vector: unconditional 16-byte load
                    L.append(f"{ind}  {_vec_load(ct, V, regs, qa, ro)}")
                elif fast == "pure":
                    # no partially valid vector in this warp: one predicated 16-byte load
                    L.append(f"{ind}  if ({g}full_{c}) {_vec_load(ct, V, regs, qa, ro)}")
                elif fast == "fastall":
                    L.append(f"{ind}  if (!wpart_{c}) {{")
                    L.append(f"{ind}    if ({g}full_{c}) {_vec_load(ct, V, regs, qa, ro)}")
                    L.append(f"{ind}  }} else {{")
                    L.extend(edge_lines(ind + "    ", qa))
                    L.append(f"{ind}  }}")
                else:
                    L.extend(edge_lines(ind + "  ", qa))
                pfd = int(self.opts.get("l2_prefetch", 2))
                if pfd > 0:
                    # fire-and-forget L2 prefetch of the row `pfd` march steps ahead (no registers,
                    # no scoreboard): later LDGs of this warp hit in L2 instead of waiting for HBM
                    pfa = self._row_addr(c, pfd)
                    if fast == "interior":
                        L.append(f"{ind}  if (pf_lane) b200::prefetch_l2({pfa});")
                    else:
                        L.append(f"{ind}  if (pfl_{c} && tt < pfe_{c}) b200::prefetch_l2({pfa});")
                if not self.SJ:
                    L.append(f"{ind}  q_{c} += sj_{c};")
                L.append(f"{ind}}}")
                continue
            if fast == "igeneral":
                # interior warp outside the steady state: only the (warp-uniform) row range decides
                pitch = "SJ" if self.SJ else f"sj_{c}"
                L.append(f"{ind}{{ const int R = tt - ({v.lag});")
                L.append(f"{ind}  if (R >= T_lo + ({v.nj[0]}) && R < T_hi + ({v.nj[1]})) {_vec_load(ct, V, regs, f'p_{c} + (long long)R * {pitch} + g0', ro)}")
                L.append(f"{ind}}}")
                continue
            L.append(f"{ind}{{ const int R = tt - ({v.lag});")
            L.append(f"{ind}  if (R >= T_lo + ({v.nj[0]}) && R < T_hi + ({v.nj[1]}) && R >= rlo_{c} && R < rhi_{c} && any_{c}) {{")
            L.append(f"{ind}    const {ct}* q = p_{c} + (long long)R * sj_{c};")
            L.append(f"{ind}    if (full_{c}) {{")
            L.append(f"{ind}      {_vec_load(ct, V, regs, 'q + g0', ro)}")
            L.append(f"{ind}    }} else {{")
            for e in range(V):
                L.append(
                    f"{ind}      if (g0 + {e} >= clo_{c} && g0 + {e} < chi_{c}) {regs[e]} = {ldf}(q + (long long)(g0 + {e}) * si_{c});"
                )
            L.append(f"{ind}    }}")
            L.append(f"{ind}  }}")
            L.append(f"{ind}}}")
        # 2. stages
        for si, he in enumerate(self.hes):
            L.extend(self._emit_stage(si, he, phase, fast))
        return L

    def _emit_ring_loads(self, phase: int, u: int) -> List[str]:
        """Interior steady step of the bulk-async variant: this lane's vector of every input row of step `u` of the
        trip, read from the trip's slot in shared memory (the slot's mbarrier was waited for at the top of the trip)."""
        V, RB = self.V, self.tma_rb
        ins = [v for v in self.values if v.kind == "in"]
        ind = "      "
        L = []
        for n, v in enumerate(ins):
            regs = [self.reg(v, self._new_slot(v, phase), e) for e in range(V)]
            L.append(f"{ind}{_vec_load(CT[v.dtype], V, regs, f'(rp + {(n * self.tma_ut + u) * RB})', ro=False)}")
        if not self.SJ:
            for v in ins:
                L.append(f"{ind}q_{v.cname} += sj_{v.cname};")
        return L

    def _row_addr(self, c: str, ahead: int = 0) -> str:
        """Address of this lane's vector in the row the current steady step touches (+ `ahead` rows)."""
        if self.SJ and not self._useq:
            k = self._u + ahead - self._lag_of[c]
            return f"(p_{c} + toff + ({k}) * SJ)"
        if self.SJ:
            k = self._u + ahead
            return f"(q_{c} + {k} * SJ)" if k else f"q_{c}"
        return f"(q_{c} + {ahead} * sj_{c})" if ahead else f"q_{c}"

    def _emit_stage(self, si: int, he: dict, phase: int, fast) -> List[str]:
        V = self.V
        ind = "      "
        L: List[str] = [f"{ind}{{  // stage {si}, row r = tt - {self.lag[si]}"]
        ind2 = ind + "  "
        L.append(f"{ind2}const int r = tt - ({self.lag[si]});")
        shuf_defined: set = set()
        mask_counter = [0]
        kern = self

        class EG(cg.ExprGen):
            def __init__(self, elem):
                super().__init__(kern.ft, set(), args="A")
                self.elem = elem
                self.pre: List[str] = []
                if kern.opts.get("div_inv", True):
                    self.div_hoist = lambda cx: kern.div_hoisted.setdefault(cx, f"dv{len(kern.div_hoisted)}")
                    self.div_group = self._div_group
                self.k = "k"
                self.i = f"(g0 + {elem})"
                self.j = "r"

            def field_load(self, node):
                v, di, dj = kern.binding[id(node)]
                if v is None:
                    return f"(({CT[node['dtype']]})0)"
                if v == "direct":
                    return cg.ExprGen.field_load(self, node)
                slot = kern._slot_of(v, kern.lag[si], dj, phase)
                e = self.elem + di
                if 0 <= e < V:
                    return kern.reg(v, slot, e)
                dl, src = divmod(e, V)  # lane distance (floor) and element inside that lane's vector
                fn, tag = ("__shfl_up_sync", f"L{-dl}") if dl < 0 else ("__shfl_down_sync", f"R{dl}")
                nm = f"{kern.reg(v, slot, src)}_{tag}"
                if nm not in shuf_defined:
                    shuf_defined.add(nm)
                    self.pre.append(f"const {CT[v.dtype]} {nm} = {fn}(0xffffffffu, {kern.reg(v, slot, src)}, {abs(dl)});")
                return nm

            def _div_group(self, a: str, dv: str, ct: str) -> str:
                # quotients by a launch-invariant divisor: computed ahead of the statement (like the shuffles); the
                # statement's whole group — all V elements — shares one range check (flush_pre)
                nm = f"dq{len(divq)}"
                acc = f"dbad{divg[0]}{'f' if ct == 'float' else 'd'}"
                # a dividend that contains an earlier quotient of the group (x / dx / dy) must be re-evaluated from the
                # redone quotient when the group falls back to IEEE divisions
                redo = a if any(q[0] in a for q in divq[divg[1]:]) else None
                divq.append((nm, dv, ct, acc, redo))
                self.pre.append(f"{ct} {nm}_a = {a}; {ct} {nm} = b200::div_inv_try({nm}_a, {dv}, {acc});")
                return nm

            def expr(self, n):
                if n["t"] == "scalar" and n["name"] in self.locals:
                    return f"{self.locals[n['name']]}_{self.elem}"
                return super().expr(n)

        egs = [EG(e) for e in range(V)]
        for d in he["locals"]:
            for eg in egs:
                eg.locals[d["name"]] = f"l_{d['name']}"
            ct = CT[d["dtype"]]
            L.append(ind2 + f"{ct} " + ", ".join(f"l_{d['name']}_{e} = ({ct})0" for e in range(V)) + ";")

        divq: List[tuple] = []  # quotients of the stage so far: (name, DivInv, C type, accumulator, dividend to re-evaluate or None)
        divg = [0, 0]  # [current group index, first quotient of the current group]

        def flush_pre():
            group = divq[divg[1]:]
            for acc, ct in sorted({(q[3], q[2]) for q in group}):
                L.append(ind2 + f"{'unsigned' if ct == 'float' else 'unsigned long long'} {acc} = 0;")
            for eg in egs:
                for line in eg.pre:
                    L.append(ind2 + line)
                eg.pre = []
            if group:
                # one (warp-uniform) branch for the group: some quotient left the guarded exponent range (zero, infinity,
                # NaN, tiny, huge) -> the IEEE division for all of them, the same values where the fast path was valid
                cond = " || ".join(f"b200::div_inv_bad({acc})" for acc in sorted({q[3] for q in group}))
                L.append(ind2 + f"if (__builtin_expect(__any_sync(0xffffffffu, {cond}), 0)) {{")
                L.append(ind2 + "  B200_TRACE(6);")
                for nm, dv, ct, acc, redo in group:
                    if redo is not None:
                        L.append(ind2 + f"  {nm}_a = {redo};")
                    # (out-of-line: a call per quotient keeps the loop small but its ABI costs registers once the loop is
                    #  unrolled 4+ times — 71 / 96 / 108 for upwind5 at period 2 / 4 / 8 — hence inline there)
                    inline = kern.opts.get("div_slow", "call" if kern.U <= 2 else "inline") == "inline"
                    L.append(ind2 + (f"  {nm} = {nm}_a / {dv}.d;" if inline else f"  {nm} = b200::div_ieee({nm}_a, {dv}.d);"))
                L.append(ind2 + "}")
                divg[0] += 1
                divg[1] = len(divq)

        def emit_stmts(stmts, masks: List[Optional[str]]):
            # masks: per-element condition variable names (None = unconditional)
            for s in stmts:
                t = s["t"]
                if t == "assign":
                    rhs = [egs[e].expr(s["right"]) for e in range(V)]
                    flush_pre()
                    left = s["left"]
                    if left["t"] == "scalar":
                        ct = CT[left["dtype"]]
                        for e in range(V):
                            tgt = f"l_{left['name']}_{e}"
                            val = f"({ct})({rhs[e]})"
                            L.append(ind2 + (f"{tgt} = {masks[e]} ? {val} : {tgt};" if masks[e] else f"{tgt} = {val};"))
                    else:
                        new, prev = kern.assign_ver[id(s)]
                        ct = CT[new.dtype]
                        slot = kern._new_slot(new, phase)
                        for e in range(V):
                            tgt = kern.reg(new, slot, e)
                            val = f"({ct})({rhs[e]})"
                            if masks[e]:
                                if prev is None:
                                    old = f"({ct})0"
                                else:
                                    old = kern.reg(prev, kern._slot_of(prev, kern.lag[si], 0, phase), e)
                                L.append(ind2 + f"{tgt} = {masks[e]} ? {val} : {old};")
                            else:
                                L.append(ind2 + f"{tgt} = {val};")
                        if new.store:
                            L.extend(kern._emit_store(new, slot, ind2, fast))
                elif t == "mask":
                    conds = [egs[e].expr(s["mask"]) for e in range(V)]
                    flush_pre()
                    mid = mask_counter[0]
                    mask_counter[0] += 1
                    names = []
                    for e in range(V):
                        nm = f"m{mid}_{e}"
                        c = f"({conds[e]})" + (f" && {masks[e]}" if masks[e] else "")
                        L.append(ind2 + f"const bool {nm} = {c};")
                        names.append(nm)
                    emit_stmts(s["body"], names)
                elif t == "hregion":
                    mid = mask_counter[0]
                    mask_counter[0] += 1
                    names = []
                    for e in range(V):
                        conds = []
                        for var, n_sym, (lo, hi) in ((f"(g0 + {e})", "A.g.nI", s["i"]), ("r", "A.g.nJ", s["j"])):
                            if lo is not None:
                                conds.append(f"{var} >= {cg._bound(lo, n_sym)}")
                            if hi is not None:
                                conds.append(f"{var} < {cg._bound(hi, n_sym)}")
                        if masks[e]:
                            conds.append(masks[e])
                        nm = f"m{mid}_{e}"
                        L.append(ind2 + f"const bool {nm} = {' && '.join(conds) if conds else 'true'};")
                        names.append(nm)
                    emit_stmts(s["body"], names)
                else:  # pragma: no cover
                    raise NotStreamable(t)

        emit_stmts(he["body"], [None] * V)
        L.append(f"{ind}}}")
        return L

    def _emit_store(self, v: Value, slot: int, ind: str, fast) -> List[str]:
        V = self.V
        c = v.cname
        e = self.hes[v.stage]["extent"]
        regs = [self.reg(v, slot, x) for x in range(V)]
        ct = CT[v.dtype]
        rowg = f"r >= A.g.j_lo + ({e[1][0]}) && r < A.g.j_hi + ({e[1][1]})"
        if fast == "igeneral":
            pitch = "SJ" if self.SJ else f"sj_{c}"
            return [f"{ind}if (own && r >= T_lo && r < T_hi) {_vec_store(ct, V, regs, f'p_{c} + (long long)r * {pitch} + g0')}"]
        if fast:
            g = f"{rowg} && " if self.row_guard[c] else ""
            qa = self._row_addr(c)
            if fast == "interior":
                return [f"{ind}if (own) {_vec_store(ct, V, regs, qa)}"] + ([] if self.SJ else [f"{ind}q_{c} += sj_{c};"])
            L = [f"{ind}if ({g}sfull_{c}) {_vec_store(ct, V, regs, qa)}"]
            for x in range(V if fast != "pure" else 0):
                L.append(f"{ind}if ({g}pe_{c}_{x}) ({qa})[{x}] = {regs[x]};")
            if not self.SJ:
                L.append(f"{ind}q_{c} += sj_{c};")
            return L
        L = [
            f"{ind}if (own && r >= T_lo && r < T_hi && {rowg} && any_{c}) {{",
            f"{ind}  {ct}* q = p_{c} + (long long)r * sj_{c};",
            f"{ind}  if (full_{c}) {{",
            f"{ind}    {_vec_store(ct, V, regs, 'q + g0')}",
            f"{ind}  }} else {{",
        ]
        for x in range(V):
            L.append(f"{ind}    if (g0 + {x} >= clo_{c} && g0 + {x} < chi_{c}) q[(long long)(g0 + {x}) * si_{c}] = {regs[x]};")
        L += [f"{ind}  }}", f"{ind}}}"]
        return L


def _vec_type(ct: str, V: int) -> Tuple[str, List[str]]:
    size = {"float": 4, "int": 4, "double": 8, "long long": 8}[ct] * V
    if size == 16:
        return "int4", ["x", "y", "z", "w"]
    if size == 8:
        return "int2", ["x", "y"]
    raise NotStreamable("vector size")


#: cache hints of the vector accesses of the kernel being emitted (set by StreamKernel.emit from the
#: options `ldcs` / `stcs`): streaming (evict-first) loads instead of the read-only path, streaming stores
_HINTS = {"ld": "__ldg", "st": ""}


def _vec_load(ct: str, V: int, regs: List[str], addr: str, ro: bool = True) -> str:
    """16-byte load into V scalar registers (bit casts keep any 4/8-byte element type);
    `ro`: the field is never written by this kernel -> read-only (non-coherent) path."""
    ld = _HINTS["ld"] if ro else "*"
    if ct in ("float", "int"):
        cast = "__int_as_float" if ct == "float" else ""
        comps = ["x", "y", "z", "w"][:V]
        body = " ".join(f"{regs[e]} = {cast}(u.{comps[e]});" for e in range(V))
        vt = "int4" if V == 4 else "int2"
        return f"{{ const {vt} u = {ld}(reinterpret_cast<const {vt}*>({addr})); {body} }}"
    # 8-byte elements, V == 2
    if ct == "double":
        return (
            f"{{ const int4 u = {ld}(reinterpret_cast<const int4*>({addr})); "
            f"{regs[0]} = __hiloint2double(u.y, u.x); {regs[1]} = __hiloint2double(u.w, u.z); }}"
        )
    return (
        f"{{ const longlong2 u = {ld}(reinterpret_cast<const longlong2*>({addr})); {regs[0]} = u.x; {regs[1]} = u.y; }}"
    )


def _vec_store(ct: str, V: int, regs: List[str], addr: str) -> str:
    if ct == "float":
        vt = "float4" if V == 4 else "float2"
    elif ct == "int":
        vt = "int4" if V == 4 else "int2"
    elif ct == "double":
        vt = "double2"
    else:
        vt = "longlong2"
    val = f"make_{vt}({', '.join(regs[: 2 if vt in ('double2', 'longlong2') else V])})"
    if _HINTS["st"]:
        return f"{_HINTS['st']}(reinterpret_cast<{vt}*>({addr}), {val});"
    return f"*reinterpret_cast<{vt}*>({addr}) = {val};"


# ---------------------------------------------------------------------------------------------------
# loop fusion (interval refinement) — consecutive PARALLEL computations become one set of kernels
# ---------------------------------------------------------------------------------------------------
def _bkey(b) -> Tuple[int, int]:
    return (0, int(b[1])) if b[0] == "start" else (1, int(b[1]))


def _loop_accesses(loop):
    acc = [a for sec in loop["sections"] for he in sec["hes"] for a in b2ir.field_accesses(he["body"])]
    written = {a["name"] for a in acc if a["write"]}
    reads_off = {a["name"] for a in acc if not a["write"] and (isinstance(a["off"], dict) or tuple(a["off"]) != (0, 0, 0))}
    reads_koff = {a["name"] for a in acc if not a["write"] and (isinstance(a["off"], dict) or a["off"][2] != 0)}
    return written, reads_off, reads_koff


def _can_fuse(a, b, stencil) -> bool:
    """Loop `b` may run level by level right behind loop `a` (inside one kernel per K interval) when
    no value crosses levels or threads between them through memory: b reads nothing a writes at a K
    offset (that level may not be computed yet), a reads nothing b writes at any offset (it may
    already be overwritten), and the symbolic interval bounds of both are ordered for every domain
    the stencil accepts (start+x <= end+y for all bounds, guaranteed by domain_info.min_k)."""
    wa, ra_off, _ = _loop_accesses(a)
    wb, _, rb_koff = _loop_accesses(b)
    # a zero-offset read inside a horizontal execution that runs on an extended extent is a cross-thread read too
    ra_ext = {
        acc["name"]
        for sec in a["sections"] for he in sec["hes"] if any(x != 0 for ax in he["extent"] for x in ax)
        for acc in b2ir.field_accesses(he["body"]) if not acc["write"]
    }  # fmt: skip
    if wa & rb_koff or wb & (ra_off | ra_ext):
        return False
    bounds = [bd for lp in (a, b) for sec in lp["sections"] for bd in sec["interval"]]
    starts = [bd[1] for bd in bounds if bd[0] == "start"]
    ends = [bd[1] for bd in bounds if bd[0] == "end"]
    min_k = int(stencil["domain_info"]["min_k"])
    if starts and ends and max(starts) - min(ends) > min_k:
        return False
    return True


def _fuse(a, b):
    bounds = sorted({_bkey(bd) for lp in (a, b) for sec in lp["sections"] for bd in sec["interval"]})
    unkey = lambda k: ["start" if k[0] == 0 else "end", k[1]]  # noqa: E731
    sections = []
    for lo, hi in zip(bounds, bounds[1:]):
        hes = []
        for lp in (a, b):
            for sec in lp["sections"]:
                if _bkey(sec["interval"][0]) <= lo and hi <= _bkey(sec["interval"][1]):
                    hes += sec["hes"]
        if hes:
            sections.append({"interval": [unkey(lo), unkey(hi)], "hes": hes})
    return {"order": "parallel", "sections": sections, "caches": [], "fused": True}


def fuse_parallel_loops(stencil: Dict[str, Any]) -> Dict[str, Any]:
    """Merge consecutive PARALLEL loops whenever `_can_fuse` holds: the K axis is cut at every
    interval bound of either loop and each piece runs the horizontal executions of both, so values
    handed from one computation to the next stay in the register windows of the streaming kernel
    instead of a round trip through scratch memory (the reference reaches the same effect by
    inlining the producer, gtc/passes/oir_optimizations/inlining.py OnTheFlyMerging, which this
    backend skips because recomputation at every offset costs more than a window row)."""
    out: List[dict] = []
    changed = False
    for loop in stencil["loops"]:
        if out and loop["order"] == "parallel" and out[-1]["order"] == "parallel" and _can_fuse(out[-1], loop, stencil):
            out[-1] = _fuse(out[-1], loop)
            changed = True
        else:
            out.append(loop)
    if not changed:
        return stencil
    new = dict(stencil)
    new["loops"] = out
    return new


# ---------------------------------------------------------------------------------------------------
# driver
# ---------------------------------------------------------------------------------------------------
def _fields_touched(hes) -> set:
    return {a["name"] for he in hes for a in b2ir.field_accesses(he["body"])}


def _defined_before_use(hes, name: str) -> bool:
    """Within this section, `name` is assigned unconditionally before any read and only read at K
    offset 0: nothing flows into the section through it."""
    defined = False

    def reads(node, found):
        b2ir.walk_exprs(node, lambda e: e["t"] == "field" and e["name"] == name and found.append(e))

    def visit(stmts, top):
        nonlocal defined
        for st in stmts:
            t = st["t"]
            if t == "assign":
                found: list = []
                reads(st["right"], found)
                left = st["left"]
                if left["t"] == "field":
                    for ix in left.get("data_index", []):
                        reads(ix, found)
                for e in found:
                    if not defined or isinstance(e["off"], dict) or e["off"][2] != 0:
                        return False
                if left["t"] == "field" and left["name"] == name:
                    if top:
                        defined = True
                    elif not defined:
                        return False  # first definition under a mask: the old value survives
            else:
                found = []
                reads(st.get("mask", st.get("cond", [])), found)
                for e in found:
                    if not defined or isinstance(e["off"], dict) or e["off"][2] != 0:
                        return False
                if visit(st["body"], False) is False:
                    return False
        return True

    for he in hes:
        if visit(he["body"], True) is False:
            return False
    return True


def _generate(stencil: Dict[str, Any], options: Dict[str, Any], *, strict: bool):
    gen = cg.Generator(stencil, options)
    used_stream = False
    # fields touched per (loop, section) to decide which temporaries are kernel-local
    secs = [(li, si) for li, loop in enumerate(stencil["loops"]) for si, _ in enumerate(loop["sections"])]
    touched = {(li, si): _fields_touched(stencil["loops"][li]["sections"][si]["hes"]) for li, si in secs}
    temps = {t["name"] for t in stencil["temporaries"]}
    # a temporary that every section (re)defines before using it never carries a value between kernels
    private = {
        t for t in temps
        if all(_defined_before_use(stencil["loops"][li]["sections"][si]["hes"], t) for (li, si) in secs if t in touched[(li, si)])
        and all(stencil["loops"][li]["order"] == "parallel" for (li, si) in secs if t in touched[(li, si)])
    }  # fmt: skip
    pending: List[dict] = []  # run of consecutive sequential loops (may fuse into one column kernel)
    for li, loop in enumerate(stencil["loops"]):
        if loop["order"] != "parallel":
            pending.append(loop)
            continue
        gen.lower_loops(pending)
        pending = []
        plans = []
        ok = True
        for si, sec in enumerate(loop["sections"]):
            others = set().union(*[f for key, f in touched.items() if key != (li, si)]) if len(touched) > 1 else set()
            global_fields = (temps & others) - private
            sk = StreamKernel(gen, sec["interval"], sec["hes"], global_fields, options)
            try:
                sk.analyse()
            except NotStreamable:
                ok = False
                break
            plans.append(sk)
        if not ok:
            if strict and loop.get("fused"):
                return None  # a fused loop must stream; otherwise generate from the unfused stencil
            gen.lower_loop(loop)
            continue
        for sk in plans:
            k = sk.emit()
            gen.steps.append({"t": "launch", "kernel": k})
            used_stream = True
    gen.lower_loops(pending)
    if not used_stream:
        return None
    return gen.finish()


def try_generate(stencil: Dict[str, Any], options: Dict[str, Any]):
    """Generate with streaming kernels for every PARALLEL section that fits the template; other
    loops use the point generator.  Returns None when nothing is streamable."""
    if options.get("fuse_loops", True):
        fused = fuse_parallel_loops(stencil)
        if fused is not stencil:
            res = _generate(fused, options, strict=True)
            if res is not None:
                return res
    return _generate(stencil, options, strict=False)
