"""Differential fuzzing of the b200 code generators (dev tool; build container only: needs the gt4py
frontend from /root/reference through tools/shims, and g++ for the CPU emulator).

Random GTScript stencils -> gt4py frontend + OIR passes -> b200 IR -> (a) the NumPy oracle and
(b) the GENERATED CUDA kernels executed by tests/emu (C-order arrays and the backend's storage layout
between guard pages), compared bit for bit.  Two families:

  par   several PARALLEL computations with vertical intervals, temporaries read at IJ offsets, inputs
        read at IJK offsets, ternaries / if-else blocks, min/max/abs  (streaming generator, loop fusion)
  col   FORWARD + BACKWARD sweeps with interval splits and k-1 / k+1 reads of swept fields, masked
        updates  (column generator: carried registers, prefetch, deferred stores)

    PYTHONPATH=tools/shims:/root/reference/src:.:tests GT_CACHE_ROOT=/tmp/gtcache python tools/fuzz_codegen.py --n 200 --seed 0
"""

from __future__ import annotations

import argparse
import importlib.util
import pathlib
import random
import sys
import tempfile
import warnings

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

HEADER = """import numpy as np
from gt4py.cartesian.gtscript import PARALLEL, FORWARD, BACKWARD, Field, computation, interval, horizontal, region, I, J, K
F = Field[np.{dtype}]
"""


class Gen:
    def __init__(self, rng: random.Random, dtype: str):
        self.r = rng
        self.dtype = dtype

    def lit(self) -> str:
        return self.r.choice(["0.5", "2.0", "0.25", "1.5", "3.0", "0.125", "1.0"])

    def ij(self, reach: int):
        return self.r.randint(-reach, reach), self.r.randint(-reach, reach)

    def leaf(self, inputs, temps, *, koff: bool, reach: int = 2) -> str:
        pool = [("in", n) for n in inputs] + [("tmp", n) for n in temps]
        kind, name = self.r.choice(pool) if pool else ("lit", "")
        if kind == "lit" or self.r.random() < 0.12:
            return self.lit()
        di, dj = self.ij(reach) if self.r.random() < 0.6 else (0, 0)
        dk = self.r.choice([-1, 0, 0, 1]) if (kind == "in" and koff) else 0
        return f"{name}[{di}, {dj}, {dk}]"

    def expr(self, inputs, temps, depth: int, **kw) -> str:
        if depth <= 0 or self.r.random() < 0.25:
            return self.leaf(inputs, temps, **kw)
        c = self.r.random()
        a = self.expr(inputs, temps, depth - 1, **kw)
        b = self.expr(inputs, temps, depth - 1, **kw)
        if c < 0.55:
            return f"({a} {self.r.choice(['+', '-', '*'])} {b})"
        if c < 0.65:
            return f"{self.r.choice(['min', 'max'])}({a}, {b})"
        if c < 0.72:
            return f"abs({a})"
        if c < 0.78:
            return f"({a} / (abs({b}) + 1.0))"
        if c < 0.84:  # launch-invariant divisor: the hoisted-reciprocal division of the generators (b200::DivInv)
            return f"({a} / {self.r.choice(['3.0', '0.75', '60.0', '7.0', '1.0e-3', '(2.0 * 3.0)'])})"
        cond = f"({self.leaf(inputs, temps, **kw)} {self.r.choice(['>', '<', '>=', '<='])} {self.leaf(inputs, temps, **kw)})"
        return f"({a} if {cond} else {b})"

    # ---- PARALLEL family -------------------------------------------------------------------------
    def par(self, name: str) -> str:
        r = self.r
        inputs = ["a", "b", "c"][: r.randint(1, 3)]
        outs = ["o1", "o2"][: r.randint(1, 2)]
        L = [f"def {name}({', '.join(f'{n}: F' for n in inputs + outs)}):"]
        temps: list = []
        n_comp = r.randint(1, 4)
        # K-offset reads of inputs make K halos necessary in every interval: keep them to full-interval computations
        for ci in range(n_comp):
            split = r.choice([None, None, "lo", "hi"]) if ci > 0 or n_comp == 1 else None
            ivs = {None: ["..."], "lo": ["0, 1", "1, None"], "hi": ["0, -1", "-1, None"]}[split]
            L.append("    with computation(PARALLEL):")
            new_t = [f"t{len(temps) + i}" for i in range(r.randint(1, 3))] if ci < n_comp - 1 else []
            for iv in ivs:
                L.append(f"        with interval({iv}):")
                body_t = list(temps)
                for t in new_t:  # temporaries must be defined in every interval piece of their computation
                    L.append(f"            {t} = {self.expr(inputs, body_t, r.randint(1, 3), koff=split is None, reach=2)}")
                    if r.random() < 0.3:  # masked redefinition
                        L.append(f"            if {self.leaf(inputs, body_t, koff=False, reach=0)} > {self.lit()}:")
                        L.append(f"                {t} = {self.expr(inputs, body_t, 2, koff=False, reach=1)}")
                    body_t.append(t)  # later statements of the same computation may read it at IJ offsets (stages)
                if ci == n_comp - 1:
                    if r.random() < 0.4 and temps:
                        cond = f"{self.leaf(inputs, temps, koff=False, reach=0)} > {self.lit()}"
                        L.append(f"            if {cond}:")
                        for o in outs:
                            L.append(f"                {o} = {self.expr(inputs, temps, 2, koff=False, reach=2)}")
                        L.append("            else:")
                        for o in outs:
                            L.append(f"                {o} = {self.expr(inputs, temps, 2, koff=False, reach=1)}")
                    else:
                        for o in outs:
                            L.append(f"            {o} = {self.expr(inputs, body_t, r.randint(1, 3), koff=split is None, reach=2)}")
                    if r.random() < 0.3:  # horizontal region on an output (positional mask, may reach outside the domain)
                        reg = r.choice(["I[0], :", ":, J[-1]", "I[0] : I[0] + 2, J[0] + 1 :", "I[-1] - 1 :, :", ":, J[0] - 1 : J[0] + 1"])
                        L.append(f"            with horizontal(region[{reg}]):")
                        L.append(f"                {outs[0]} = {self.expr(inputs, body_t, 2, koff=False, reach=1)}")
            temps += new_t
        return "\n".join(L) + "\n"

    # ---- column family -----------------------------------------------------------------------------
    def col(self, name: str) -> str:
        r = self.r
        inputs = ["a", "b"][: r.randint(1, 2)]
        swept = ["x", "y"][: r.randint(1, 2)]
        L = [f"def {name}({', '.join(f'{n}: F' for n in inputs + swept)}):"]

        def sweep(order: str, back: int):
            lo, hi = ("0, 1", "1, None") if order == "FORWARD" else ("-1, None", "0, -1")
            L.append(f"    with computation({order}):")
            L.append(f"        with interval({lo}):")
            for s in swept:
                L.append(f"            {s} = {self.expr(inputs, [], 1, koff=False, reach=0)}")
            L.append(f"        with interval({hi}):")
            known = []
            for s in swept:
                prev = [f"{q}[0, 0, {back}]" for q in swept] + [f"{q}[0, 0, 0]" for q in known]
                terms = [self.expr(inputs, [], 1, koff=False, reach=1), r.choice(prev)]
                if r.random() < 0.5:
                    terms.append(f"{r.choice(prev)} * {self.lit()}")
                L.append(f"            {s} = {' + '.join(terms)}")
                if r.random() < 0.4:
                    L.append(f"            if {s} > {self.lit()}:")
                    L.append(f"                {s} = {s} * 0.5 - {r.choice(prev)}")
                known.append(s)

        sweep("FORWARD", -1)
        if r.random() < 0.7:
            sweep("BACKWARD", 1)
        return "\n".join(L) + "\n"


def run_case(source: str, fname: str, dtype: str, seed: int, workdir: pathlib.Path, build_opts: dict) -> str:
    """-> "" when oracle == emulated kernels for every written field, else a description."""
    from emu.emu import EmuStencil
    from gt4py_b200 import from_oir, testing
    from oracle import numpy_oracle

    path = workdir / f"{fname}.py"
    path.write_text(HEADER.format(dtype=dtype) + "\n\n" + source)
    spec = importlib.util.spec_from_file_location(fname, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    try:
        irs = {v: from_oir.lower_definition(getattr(mod, fname), name=fname, variant=v, **build_opts) for v in ("staged", "default")}
    except Exception as exc:  # the generator produced something GTScript does not accept: not a finding
        return f"SKIP frontend: {type(exc).__name__}: {str(exc)[:120]}"
    rng = random.Random(seed)
    # the oracle itself against the REFERENCE numpy backend on this stencil (pins the oracle beyond the fixtures)
    from gt4py.cartesian import gtscript

    ref_stencil = gtscript.stencil(backend="numpy", definition=getattr(mod, fname), name=fname + "_ref", **build_opts)
    st0 = irs["default"]
    fields, params, origins, domain = testing.make_case_data(st0, fname, domain=(17, 6, max(4, int(st0["domain_info"]["min_k"]))), seed=seed)
    a = {k: v.copy() for k, v in fields.items() if v is not None}
    b = {k: v.copy() for k, v in fields.items() if v is not None}
    import inspect

    unused = {n: None for n in inspect.signature(getattr(mod, fname)).parameters if n not in a}  # pruned arguments
    try:
        ref_stencil(**a, **unused, **params, origin=origins, domain=domain)
    except Exception as exc:  # e.g. the reference's numpy code generator mis-shapes literal-only masks
        return f"SKIP reference numpy backend failed: {type(exc).__name__}: {str(exc)[:100]}"
    numpy_oracle.run(st0, b, params, domain, origins)
    for w in a:
        if not np.array_equal(a[w], b[w], equal_nan=True):
            return f"ORACLE != REFERENCE numpy backend, field {w}"
    for variant, st in irs.items():
        domain = (rng.choice([5, 33, 67, 130]), rng.choice([3, 9, 70]), max(rng.choice([2, 3, 5]), int(st["domain_info"]["min_k"])))
        fields, params, origins, domain = testing.make_case_data(st, fname, domain=domain, seed=seed)
        if seed % 2:  # special values: the guarded fast paths of the generators (hoisted-reciprocal division) must fall back
            srng = np.random.default_rng(seed)
            for arr in fields.values():
                if arr is not None and arr.dtype.kind == "f":
                    flat = arr.reshape(-1)
                    idx = srng.permutation(flat.size)
                    m = max(1, flat.size // 12)
                    for j, val in enumerate((0.0, -0.0, np.inf, -np.inf, np.nan, 1e-44 if arr.dtype == np.float32 else 5e-324)):
                        flat[idx[j * m:(j + 1) * m]] = val
                    flat[idx[6 * m:7 * m]] *= arr.dtype.type(1e30)
        ref = {k: (v.copy() if v is not None else None) for k, v in fields.items()}
        numpy_oracle.run(st, ref, params, domain, origins)
        pitch = {-(-v.shape[0] // 32) * 32 for v in fields.values() if v is not None}
        sp = pitch.pop() if len(pitch) == 1 else 0
        variants = [{}, {"strategy": "point"}, {"fuse_loops": False, "seq_prefetch": 2}, {"seq_cache": False},
                    {"interior_loop": True, "static_pitch": sp, "tile_j": 32},
                    {"k_order": True, "period": 4, "div_slow": "inline", "col_smem": True, "seq_prefetch": 4},
                    {"k_order": False, "interior_loop": True, "static_pitch": sp, "tma": 2, "tile_j": 16, "seq_rotate": True, "seq_prefetch": 2, "fuse_columns": True},
                    {"div_inv": False, "seq_rotate": False, "col_hints": True}]  # fmt: skip
        for opts in variants:
            es = EmuStencil(st, opts, name=f"{fname}.{variant}")
            for layout in (None, "b200"):
                got = {k: (v.copy() if v is not None else None) for k, v in fields.items()}
                es.run(got, params, domain, origins, layout=layout, guard="end")
                for w in st["field_info"]:
                    if got.get(w) is None:
                        continue
                    if not np.array_equal(got[w], ref[w], equal_nan=True):
                        bad = np.argwhere(~((got[w] == ref[w]) | (np.isnan(got[w]) & np.isnan(ref[w]))))
                        return f"MISMATCH {variant} {opts} layout={layout} field {w} domain {domain}: {len(bad)} cells, first {bad[:3].tolist()}"
    return ""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=50)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--family", default="both", choices=["par", "col", "both"])
    ap.add_argument("--keep", default=None, help="directory for the generated sources (default: a temp dir)")
    args = ap.parse_args()
    warnings.filterwarnings("ignore")
    workdir = pathlib.Path(args.keep or tempfile.mkdtemp(prefix="b200_fuzz_"))
    workdir.mkdir(parents=True, exist_ok=True)
    stats = {"ok": 0, "skip": 0, "fail": 0}
    for n in range(args.n):
        seed = args.seed * 100003 + n
        rng = random.Random(seed)
        family = args.family if args.family != "both" else rng.choice(["par", "col"])
        dtype = rng.choice(["float32", "float64"])
        build = {"literal_float_precision": 32} if (dtype == "float32" and rng.random() < 0.7) else {}
        fname = f"fz_{family}_{seed}"
        gen = Gen(rng, dtype)
        source = gen.par(fname) if family == "par" else gen.col(fname)
        res = run_case(source, fname, dtype, seed, workdir, build)
        if not res:
            stats["ok"] += 1
        elif res.startswith("SKIP"):
            stats["skip"] += 1
            print(f"[{n}] {fname}: {res}", flush=True)
        else:
            stats["fail"] += 1
            print(f"[{n}] {fname}: {res}\n{source}", flush=True)
    print(f"fuzz: {stats} (sources in {workdir})")
    sys.exit(1 if stats["fail"] else 0)


if __name__ == "__main__":
    main()
