"""Column ("K-march") generator for FORWARD / BACKWARD computation blocks.

One thread owns one (i, j) column (I-coalesced across the warp) and marches along K inside the
thread, like the baseline `seq` kernel of `codegen.py`, but with the K data flow of the column kept
in registers instead of going back to memory for every statement:

* **k-caches in registers** — a value read at a constant offset `f[di, dj, dk]` that the column has
  already touched one level earlier (written, or read at `dk + step`) is *carried* in a register and
  rotated at the end of the level (`sup[0,0,-1]`, `rhs[0,0,-1]` of the Thomas forward sweep,
  `out[0,0,1]` of the back substitution): every cell is loaded at most once per sweep.
* **look-ahead prefetch** (`seq_prefetch` = depth, default 1 level) — every remaining unconditional
  load of the level (`inf`, `diag`, the not-yet-updated `sup`/`rhs`) is issued `depth` levels early
  into a register pipeline, at the top of an earlier level's body, so the loads of level k+1 are in flight while the (long, dependent)
  fp64 division chain of level k executes.  Little's law then holds with the ~55 resident warps/SM
  a 512x512 plane provides, without relying on the compiler to hoist loads over possibly-aliasing
  stores (it cannot: every field is a `double*` out of the same argument block).
* stores are never deferred (memory is always current), so variable-K reads, later sections and
  later kernels need no flush logic; masks stay real branches (lazy evaluation like the baseline).

This is the b200 counterpart of GridTools' k-caches with fill/flush
(reference: gtc/gtcpp/gtcpp_codegen.py:227-247, gtc/passes/oir_optimizations/caches.py) — the
reference's cache annotations are not used, the analysis below is exact for the generated code.

Anything outside the template (while loops, writes at an offset, horizontal executions with
different extents, …) raises `NotColumnable` and the caller falls back to the baseline kernel.
Semantics: SURVEY.md §9 (numpy backend: `for k_ in range(k, K)` / reversed, statement by statement).
"""

from __future__ import annotations

from typing import Any, Dict, List, Optional, Set, Tuple

from . import ir as b2ir

CT = b2ir.CTYPE

Key = Tuple[str, int, int, int]  # (field, di, dj, dk)


class NotColumnable(Exception):
    pass


def _tag(key: Key) -> str:
    def n(v):
        return f"m{-v}" if v < 0 else f"p{v}"

    name, di, dj, dk = key
    return f"{''.join(c if c.isalnum() else '_' for c in name)}_{n(di)}{n(dj)}{n(dk)}"


class ColumnKernel:
    RING_DEPTH = 3  # default look-ahead (levels) of the register ring

    def __init__(self, gen, loops: List[dict], opts: Dict[str, Any], external: Optional[Set[str]] = None):
        from . import codegen as cg

        self.cg = cg
        self.gen = gen
        self.ft = gen.ft
        self.loops = loops  # consecutive sweeps executed back to back by the same thread (see fusable())
        self.opts = opts
        self.step = 1
        self.hes = [he for loop in loops for sec in loop["sections"] for he in sec["hes"]]
        if not self.hes:
            raise NotColumnable("empty loop")
        ext = self.hes[0]["extent"]
        if any(he["extent"] != ext for he in self.hes):
            raise NotColumnable("horizontal executions with different extents")
        self.extent = ext
        acc = [a for he in self.hes for a in b2ir.field_accesses(he["body"])]
        self.written: Set[str] = {a["name"] for a in acc if a["write"]}
        self.read: Set[str] = {a["name"] for a in acc if not a["write"]}
        self.loop_written: Set[str] = set(self.written)  # fields written by the sweep being emitted
        for a in acc:
            if a["write"] and (isinstance(a["off"], dict) or tuple(a["off"]) != (0, 0, 0)):
                raise NotColumnable("write at an offset")
        self._check_stmts([s for he in self.hes for s in he["body"]])
        # fields whose constant-offset accesses are kept in registers: plain IJK fields
        self.cacheable: Set[str] = set()
        for a in acc:
            d = self.ft.entries[self.ft.index[a["name"]]]
            if all(d["dims"]) and not d["data_dims"]:
                self.cacheable.add(a["name"])
        # fields also read straight from memory (variable / absolute K index): their stores are
        # never deferred, so those reads always see the latest value
        self.direct_read: Set[str] = {a["name"] for a in acc if not a["write"] and isinstance(a["off"], dict)}
        self.div_hoisted: Dict[str, str] = {}
        self.smem_fields: List[str] = self._smem_candidates(acc, external) if len(loops) > 1 and opts.get("col_smem", False) else []
        self.smem_mode = False
        self._slot: Optional[int] = None
        self._ring: Set[Key] = set()
        pf = opts.get("seq_prefetch", True)
        #: levels of load look-ahead (explicit), or None: 1 for the shifting pipeline, RING_DEPTH for the register ring
        self.depth_opt: Optional[int] = None if pf is True else (0 if pf is False else max(0, int(pf)))
        self.depth = 1 if self.depth_opt is None else self.depth_opt
        self.prefetch = self.depth > 0

    def _smem_candidates(self, acc, external: Optional[Set[str]]) -> List[str]:
        """Temporaries of fused sweeps that can live in SHARED memory instead of global scratch (`col_smem`): the
        forward sweep of an implicit vertical solve hands its coefficients to the back substitution through temporaries
        (`ccol`, `dcol`); when both sweeps run in the same thread and nothing else touches those temporaries, a column's
        values are private to the thread from the first write to the last read: nK x 4 (8) bytes per thread and field,
        indexed [k][thread] (conflict-free), never written to or re-read from HBM (fast-waves w solver: 48 -> 32 B/cell).
        The b200 counterpart of GridTools' k-caches without fill / flush (gtc/gtcpp/gtcpp_codegen.py:229-232).
        Conditions: a temporary no other kernel touches (`external` = fields of the loops outside this kernel), accessed
        in its own column only, at constant K offsets that stay inside [0, nK) for every domain."""
        if external is None:
            return []
        out = []
        min_k = int(self.gen.st["domain_info"]["min_k"])
        for name in sorted(self.cacheable):
            d = self.ft.entries[self.ft.index[name]]
            if d["kind"] != "temp" or name in external:
                continue
            mine = [a for a in acc if a["name"] == name]
            if any(isinstance(a["off"], dict) or (a["off"][0], a["off"][1]) != (0, 0) or a.get("data_index") for a in mine):
                continue
            ok = True
            for loop in self.loops:
                for sec in loop["sections"]:
                    (b0, o0), (b1, o1) = sec["interval"]
                    for he in sec["hes"]:
                        for a in b2ir.field_accesses(he["body"]):
                            if a["name"] != name:
                                continue
                            dk = int(a["off"][2])
                            # lowest level touched k0 + dk >= 0, highest k1 - 1 + dk <= nK - 1, for every domain the
                            # stencil accepts (nK >= domain_info.min_k)
                            if dk < 0 and not (o0 + dk >= 0 if b0 == "start" else min_k + o0 + dk >= 0):
                                ok = False
                            if dk > 0 and not (o1 + dk <= 0 if b1 == "end" else o1 + dk <= min_k):
                                ok = False
            if ok:
                out.append(name)
        # 8-byte fields first: every field's slice of the dynamic shared memory stays naturally aligned
        return sorted(out, key=lambda n: (-b2ir.ITEMSIZE[self.ft.entries[self.ft.index[n]]["dtype"]], n))

    def _check_stmts(self, stmts) -> None:
        for s in stmts:
            if s["t"] == "while":
                raise NotColumnable("while loop")
            if s["t"] in ("mask", "hregion"):
                self._check_stmts(s["body"])

    # ---- addressing ------------------------------------------------------------------------------
    def _addr(self, key: Key, kexpr: str) -> str:
        name, di, dj, dk = key
        n = self.ft.index[name]
        f = f"A.f[{n}]"
        off = ""
        if di:
            off += f" + ({di}) * {f}.s[0]"
        if dj:
            off += f" + ({dj}) * {f}.s[1]"
        kk = f"({kexpr} + ({dk}))" if dk else kexpr
        if self.smem_mode and name in self.smem_fields:
            return f"s_{self.cg._cname(name)}[{kk} * {self.nthreads}]"
        return f"c_{self.cg._cname(name)}[(long long){kk} * {f}.s[2]{off}]"

    def _load(self, key: Key, kexpr: str) -> str:
        name = key[0]
        if name in self.written:
            return self._addr(key, kexpr)
        ct = CT[self.ft.entries[self.ft.index[name]]["dtype"]]
        if self.opts.get("col_hints", False) and ct in ("float", "double", "int", "long long"):
            # read once by this kernel: streaming (evict-first) load, so that the lines the back substitution re-reads
            # (the forward sweep's own stores) are what stays in L2
            return f"__ldcs(&{self._addr(key, kexpr)})"
        return f"b200::ldro<{ct}>(&{self._addr(key, kexpr)})"

    def _store(self, key: Key, var: str) -> str:
        name = key[0]
        if self.opts.get("col_hints", False) and name not in self.read and not (self.smem_mode and name in self.smem_fields) \
                and self._ctype(key) in ("float", "double", "int", "long long"):
            return f"__stcs(&{self._addr(key, 'k')}, {var});"  # written, never read back here: streaming store
        return f"{self._addr(key, 'k')} = {var};"

    def _r(self, key: Key) -> str:
        """Register of a column value.  Prefetched keys live in a ring of D + 1 registers: the copy of the level loop
        that is being emitted (`self._slot`) consumes its own slot (no moves between the slots, see _emit_sweeps)."""
        if self._slot is not None and key in self._ring:
            return f"q{self._slot}_{_tag(key)}"
        return f"r_{_tag(key)}"

    def _ctype(self, key: Key) -> str:
        return CT[self.ft.entries[self.ft.index[key[0]]]["dtype"]]

    # ---- one level of one section: symbolic emission ---------------------------------------------
    def _emit_level(self, sec: dict, live_in: Set[Key]):
        """Emit the statements of one K level.  `live_in`: keys whose register holds the value of
        this level on entry.  Returns (lines, exposed, final): `exposed` = keys loaded at top level
        (unconditionally) because they were not valid; `final` = keys valid at the end of the level."""
        kern = self
        valid: Set[Key] = set(live_in)
        exposed: List[Key] = []
        touched: Set[Key] = set(live_in)
        L: List[str] = []
        depth = [0]

        class EG(self.cg.ExprGen):
            def __init__(self):
                super().__init__(kern.ft, kern.written, args="A")
                self.pre: List[str] = []
                self.ind = "      "
                if kern.opts.get("div_inv", True):  # divisors that are launch invariants: hoisted reciprocal (b200::DivInv)
                    self.div_hoist = lambda cx: kern.div_hoisted.setdefault(cx, f"dv{len(kern.div_hoisted)}")

            def field_load(self, node):
                off = node["off"]
                name = node["name"]
                if isinstance(off, dict) or node.get("data_index") or name not in kern.cacheable:
                    return super().field_load(node)
                key = (name, int(off[0]), int(off[1]), int(off[2]))
                var = kern._r(key)
                if key not in valid:
                    self.pre.append(f"{var} = {kern._load(key, 'k')};")
                    valid.add(key)
                    touched.add(key)
                    if depth[0] == 0:
                        exposed.append(key)
                return var

        eg = EG()

        def flush(ind):
            for line in eg.pre:
                L.append(ind + line)
            eg.pre = []

        def stmts(body, ind):
            for s in body:
                t = s["t"]
                if t == "assign":
                    rhs = eg.expr(s["right"])
                    left = s["left"]
                    if left["t"] == "scalar":
                        flush(ind)
                        L.append(f"{ind}{eg.locals[left['name']]} = ({CT[left['dtype']]})({rhs});")
                        continue
                    name = left["name"]
                    ct = CT[kern.ft.entries[kern.ft.index[name]]["dtype"]]
                    if name in kern.cacheable and not left.get("data_index"):
                        key = (name, 0, 0, 0)
                        var = kern._r(key)
                        flush(ind)
                        L.append(f"{ind}{var} = ({ct})({rhs});")
                        # an unconditional write makes every store of this level to the cell but the
                        # last one dead: keep the value in the register, store once at the end
                        if depth[0] == 0 and name not in kern.direct_read:
                            pending[key] = True
                        if key not in pending:
                            L.append(f"{ind}{kern._store(key, var)}")
                        # inside a branch a first definition is valid until the branch ends (dropped
                        # there); a key that was valid before stays valid on both paths
                        valid.add(key)
                        touched.add(key)
                    else:
                        ref = eg.field_ref(left, for_write=True)  # index expressions may load
                        flush(ind)
                        L.append(f"{ind}{ref} = ({ct})({rhs});")
                elif t in ("mask", "hregion"):
                    if t == "mask":
                        cond = eg.expr(s["mask"])
                    else:
                        conds = []
                        for var, n_sym, (lo, hi) in (("i", "A.g.nI", s["i"]), ("j", "A.g.nJ", s["j"])):
                            if lo is not None:
                                conds.append(f"{var} >= {kern.cg._bound(lo, n_sym)}")
                            if hi is not None:
                                conds.append(f"{var} < {kern.cg._bound(hi, n_sym)}")
                        cond = " && ".join(conds) if conds else "true"
                    flush(ind)
                    L.append(f"{ind}if ({cond}) {{")
                    scope_base.append(set(valid))
                    depth[0] += 1
                    stmts(s["body"], ind + "  ")
                    depth[0] -= 1
                    base = scope_base.pop()
                    # registers first defined inside the branch are undefined on the other path
                    for key in list(valid):
                        if key not in base:
                            valid.discard(key)
                    L.append(f"{ind}}}")
                else:  # pragma: no cover
                    raise NotColumnable(t)

        scope_base: List[Set[Key]] = []
        pending: Dict[Key, bool] = {}  # written cells whose store is deferred to the end of the level
        for he in sec["hes"]:
            L.append("      {")
            saved = dict(eg.locals)
            for d in he["locals"]:
                eg.locals[d["name"]] = f"l_{d['name']}"
                L.append(f"        {CT[d['dtype']]} l_{d['name']} = ({CT[d['dtype']]})0;")
            stmts(he["body"], "        ")
            L.append("      }")
            eg.locals = saved
        for key in pending:
            L.append(f"      {kern._store(key, kern._r(key))}")
        return L, exposed, valid, touched

    # ---- per-section plan ------------------------------------------------------------------------
    def _plan_section(self, sec: dict):
        step = self.step
        _, exposed, final0, _ = self._emit_level(sec, set())
        carried: List[Key] = []
        passthrough: List[Key] = []
        for key in exposed:
            name, di, dj, dk = key
            # nearest source along the march direction that is valid at the end of a level
            for n in range(1, 4):
                src = (name, di, dj, dk + n * step)
                if src in final0:
                    chain = [(name, di, dj, dk + m * step) for m in range(n)]
                    for c in chain:
                        if c not in carried:
                            carried.append(c)
                            if c != key and c not in exposed:
                                passthrough.append(c)
                    break
        prefetched: List[Key] = []
        if self.prefetch:
            for key in exposed:
                if key in carried:
                    continue
                name, di, dj, dk = key
                # the cell loaded `depth` levels early must not be written by the levels in between
                if name in self.loop_written and -self.depth <= dk * step < 0:
                    continue
                if self.smem_mode and name in self.smem_fields:
                    continue  # shared memory: no look-ahead pipeline needed
                prefetched.append(key)
        return carried, prefetched

    # ---- kernel ----------------------------------------------------------------------------------
    def emit(self) -> int:
        cg, gen = self.cg, self.gen
        name = gen._kname("col")
        bx, by = gen.BLOCK_SEQ
        if self.smem_fields:
            bx, by = (int(x) for x in self.opts.get("col_smem_block", (32, 2)))  # small CTAs: occupancy in fine steps
        self.nthreads = bx * by
        (ei0, ei1), (ej0, ej1) = self.extent
        L = [f'extern "C" __global__ void __launch_bounds__({bx * by}) {name}(const __grid_constant__ Args A) {{']
        w = L.append
        w(f"  const int i = A.g.i_lo + ({ei0}) + (int)(blockIdx.x * {bx} + threadIdx.x);")
        w(f"  const int j = A.g.j_lo + ({ej0}) + (int)(blockIdx.y * {by} + threadIdx.y);")
        w(f"  if (i >= A.g.i_hi + ({ei1}) || j >= A.g.j_hi + ({ej1})) return;")
        div_marker = len(L)
        for fname in sorted(self.cacheable):
            n = self.ft.index[fname]
            ct = CT[self.ft.entries[n]["dtype"]]
            const = "" if fname in self.written else "const "
            w(f"  {const}{ct}* const c_{cg._cname(fname)} = ({const}{ct}*)A.f[{n}].p + (long long)i * A.f[{n}].s[0] + (long long)j * A.f[{n}].s[1];")
        per_k = sum(b2ir.ITEMSIZE[self.ft.entries[self.ft.index[n]]["dtype"]] for n in self.smem_fields) * self.nthreads
        kcap = (int(self.opts.get("col_smem_kb", 56)) * 1024) // per_k if per_k else 0
        if kcap < 4:
            self.smem_fields, per_k, kcap = [], 0, 0
        self.smem_per_k, self.smem_kcap = per_k, kcap
        if self.smem_fields:
            # columns short enough for the shared-memory budget of a CTA keep these temporaries on chip (the launcher
            # requests smem_per_k x nK bytes under the same condition); taller ones take the global-scratch code below
            w("  B200_DYN_SMEM(b200_dsm);")
            w(f"  if (A.g.nK <= {kcap}) {{")
            off = 0
            for fname in self.smem_fields:
                ct = CT[self.ft.entries[self.ft.index[fname]]["dtype"]]
                w(f"    {ct}* const s_{cg._cname(fname)} = reinterpret_cast<{ct}*>(b200_dsm + (size_t)A.g.nK * {off}) + (threadIdx.y * {bx} + threadIdx.x);")
                off += b2ir.ITEMSIZE[self.ft.entries[self.ft.index[fname]]["dtype"]] * self.nthreads
            self.smem_mode = True
            self._emit_sweeps(L)
            self.smem_mode = False
            w("    return;")
            w("  }")
        self._emit_sweeps(L)
        w("}")
        L[div_marker:div_marker] = [f"  const auto {nm} = b200::div_inv_make({cx});" for cx, nm in self.div_hoisted.items()]
        gen.src.append("\n".join(L))
        gen.live |= {a["name"] for he in self.hes for a in b2ir.field_accesses(he["body"])}
        gen.kernels.append(
            {"name": name, "kind": "seq", "block": [bx, by, 1], "extent": [list(self.extent[0]), list(self.extent[1])],
             "k_lo": ["start", 0], "k_hi": ["start", 1], "smem": self._smem_pad(), "smem_per_k": per_k, "smem_kcap": kcap,
             "smem_fields": list(self.smem_fields)}
        )  # fmt: skip
        return len(gen.kernels) - 1

    def _emit_sweeps(self, L: List[str]) -> None:
        cg = self.cg
        w = L.append
        for li, si, loop, sec in [(li, si, lp, sec) for li, lp in enumerate(self.loops) for si, sec in enumerate(lp["sections"])]:
            self.step = 1 if loop["order"] == "forward" else -1
            fwd = self.step == 1
            self.loop_written = {a["name"] for sc in loop["sections"] for he in sc["hes"] for a in b2ir.field_accesses(he["body"]) if a["write"]}
            # which look-ahead pipeline (see below): planned at the ring's depth first (the hazard window of a look-ahead load
            # grows with the depth), re-planned at depth 1 when the section keeps the shifting pipeline
            rotate = self.opts.get("seq_rotate", "auto")
            self.depth = self.depth_opt if self.depth_opt is not None else self.RING_DEPTH
            carried, prefetched = self._plan_section(sec)
            if rotate == "auto":
                rotate = not any(key[0] in self.loop_written for key in prefetched)
            if not rotate and self.depth_opt is None:
                self.depth = 1
                carried, prefetched = self._plan_section(sec)
            live = set(carried) | set(prefetched)
            body, exposed, final, touched = self._emit_level(sec, live)
            k0 = cg._bound(sec["interval"][0], "A.g.nK")
            k1 = cg._bound(sec["interval"][1], "A.g.nK")
            w(f"  {{  // {loop['order']} sweep {li}, section {si}: carried {[_tag(c) for c in carried]}, prefetched {[_tag(p) for p in prefetched]}")
            w(f"    const int k0 = {k0}, k1 = {k1};")
            w("    if (k0 < k1) {")
            first = "k0" if fwd else "(k1 - 1)"

            def ahead(base: str, n: int) -> str:  # level n march steps after `base`, clamped to the section
                return f"(({base}) + {n} < k1 ? ({base}) + {n} : k1 - 1)" if fwd else f"(({base}) - {n} >= k0 ? ({base}) - {n} : k0)"

            def rotations(final_keys):
                # rotate the carried registers towards the next level (sources still hold this level)
                for key in sorted(carried, key=lambda c: c[3] * self.step):
                    src = (key[0], key[1], key[2], key[3] + self.step)
                    if src not in final_keys and src not in carried:
                        raise NotColumnable("carry chain")  # pragma: no cover
                    w(f"      {self._r(key)} = {self._r(src)};")

            # which pipeline: the ring for sweeps whose look-ahead loads are pure inputs (fast-waves w solver 1.53 -> 1.35 ms,
            # vadv 0.72 -> 0.65 ms at depth 3-4).  A sweep that updates its inputs in place (Thomas forward elimination:
            # sup, rhs) already runs at the HBM roofline of its actual traffic with one level of look-ahead, its warps
            # in step — there the ring only loosens the access pattern (0.470 -> 0.519 ms): it keeps the shifting form.
            D = self.depth
            if prefetched and rotate and D > 0:
                # look-ahead WITHOUT register moves: the prefetched values live in a ring of S = D + 1 registers and
                # the level loop is unrolled S times; copy u consumes slot u and, at its top, issues the loads of the
                # level D steps ahead into the slot the previous copy has just consumed.  (A shifting pipeline
                # `r = p1; p1 = p2; ...; p(D-1) = n` makes every level wait for the load issued ONE level earlier — the
                # move needs its source — so its depth never exceeded one: profiles/README.md r02n, the stall sites
                # of the w solver are those moves.)
                S = D + 1
                self._ring = set(prefetched)
                for key in sorted((touched | live) - self._ring):
                    w(f"      {self._ctype(key)} r_{_tag(key)};")
                for key in prefetched:
                    w(f"      {self._ctype(key)} " + ", ".join(f"q{u}_{_tag(key)}" for u in range(S)) + ";")
                for key in sorted(live - self._ring):
                    w(f"      r_{_tag(key)} = {self._load(key, first)};")
                for u in range(D):
                    for key in prefetched:
                        w(f"      q{u}_{_tag(key)} = {self._load(key, ahead(first, u) if u else first)};")
                w(f"      for (int kb = k0; kb < k1; kb += {S}) {{" if fwd else f"      for (int kb = k1 - 1; kb >= k0; kb -= {S}) {{")
                for u in range(S):
                    self._slot = u
                    body_u, _exp, final_u, _t = self._emit_level(sec, live)
                    w(f"      {{ const int k = kb {'+' if fwd else '-'} {u};   // ring slot {u}")
                    w("      if (k < k1) {" if fwd else "      if (k >= k0) {")
                    w(f"      const int kn = {ahead('k', D)};")
                    for key in prefetched:
                        w(f"      q{(u + D) % S}_{_tag(key)} = {self._load(key, 'kn')};")
                    L.extend(body_u)
                    rotations(final_u)
                    w("      }}")
                self._slot = None
                self._ring = set()
                w("      }")
                w("    }")
                w("  }")
                continue
            for key in sorted(touched | live):
                w(f"      {self._ctype(key)} r_{_tag(key)};")
            for key in sorted(live):
                w(f"      r_{_tag(key)} = {self._load(key, first)};")
            for key in prefetched:  # look-ahead pipeline: p<j> holds the value of the level j steps ahead
                for j in range(1, D):
                    w(f"      {self._ctype(key)} p{j}_{_tag(key)} = {self._load(key, ahead(first, j))};")
            w("      for (int k = k0; k < k1; ++k) {" if fwd else "      for (int k = k1 - 1; k >= k0; --k) {")
            if prefetched:
                w(f"      const int kn = {ahead('k', D)};")
                for key in prefetched:
                    w(f"      const {self._ctype(key)} n_{_tag(key)} = {self._load(key, 'kn')};")
            L.extend(body)
            rotations(final)
            for key in prefetched:
                chain = [f"r_{_tag(key)}"] + [f"p{j}_{_tag(key)}" for j in range(1, D)] + [f"n_{_tag(key)}"]
                for dst, src in zip(chain, chain[1:]):
                    w(f"      {dst} = {src};")
            w("      }")
            w("    }")
            w("  }")

    def _smem_pad(self) -> int:
        """`seq_smem_pad` = bytes of (unused) dynamic shared memory requested per CTA: caps the resident CTAs per SM
        (227 KB / pad) without touching the code.  For fused sweeps (`fuse_columns`) fewer resident columns keep a
        column's forward-sweep results in L2 until its back substitution re-reads them (2.5 KB per column for the
        Thomas solver at nK=160: 126 MB of L2 hold ~50 K columns = 10 warps per SM), trading occupancy — made up
        by a deeper `seq_prefetch` — for 72 -> 56 B/cell of HBM traffic.  To be measured."""
        pad = int(self.opts.get("seq_smem_pad", 0) or 0)
        if not 0 <= pad <= 227 * 1024:
            raise ValueError("seq_smem_pad must be between 0 and 232448 bytes")
        return pad


def fusable(a: dict, b: dict) -> bool:
    """Sweep `b` may follow sweep `a` inside the same thread (one launch, the column's freshest levels
    still in L1/L2 when `b` starts) when no value crosses columns between them: neither reads at an IJ
    offset a field the other one writes."""

    def rw(loop):
        acc = [x for sec in loop["sections"] for he in sec["hes"] for x in b2ir.field_accesses(he["body"])]
        written = {x["name"] for x in acc if x["write"]}
        off = {x["name"] for x in acc if not x["write"] and b2ir.ij_offset(x["off"]) != (0, 0)}
        return written, off

    wa, oa = rw(a)
    wb, ob = rw(b)
    return not (wa & ob) and not (wb & oa)


def try_emit(gen, loops, opts: Dict[str, Any], external: Optional[Set[str]] = None) -> Optional[int]:
    """Emit one or several consecutive FORWARD/BACKWARD loops (no level synchronisation needed) as ONE
    column kernel with register k-caches; returns the kernel index or None when the template does not apply."""
    if not opts.get("seq_cache", True):
        return None
    if isinstance(loops, dict):
        loops = [loops]
    try:
        return ColumnKernel(gen, list(loops), opts, external).emit()
    except NotColumnable:
        return None
