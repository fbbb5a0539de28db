"""gt4py OIR -> b200 stencil IR (only importable where the gt4py frontend is installed).

This is the one place where reference *objects* are touched: the OIR tree produced by gt4py's own
frontend and passes (reference: gtc/gtir_to_oir.py:50-266, gtc/passes/oir_pipeline.py:54-104) is
walked once and re-expressed as plain dicts (`ir.py`).  Extents are taken from the reference's own
analysis (gtc/passes/oir_optimizations/utils.py:250-330) so that the iteration spaces of the CUDA
kernels are exactly the ones the `numpy` backend uses (gtc/numpy/oir_to_npir.py:221-255).
"""

from __future__ import annotations

from typing import Any, Dict, Optional

from gt4py.cartesian.gtc import common, oir
from gt4py.cartesian.gtc.passes.oir_optimizations.utils import compute_extents

from . import ir as b2ir


def _dt(dtype: common.DataType) -> str:
    name = dtype.name.lower()
    if name not in b2ir.DTYPES:
        raise NotImplementedError(f"b200: unsupported dtype {dtype!r}")
    return name


def _bound(b) -> Optional[list]:
    if b is None:
        return None
    if isinstance(b, common.RuntimeAxisBound) or not isinstance(b.offset, int):
        raise NotImplementedError(
            "Runtime interval bounds (e.g. `with interval(0, field)`) is an experimental feature "
            "and not implemented for the `b200` backend."
        )
    return [str(b.level.value), int(b.offset)]


class _Lower:
    def __init__(self, stencil: oir.Stencil):
        self.stencil = stencil
        self.field_extents, self.block_extents = compute_extents(stencil)

    # -- expressions ------------------------------------------------------------------------
    def expr(self, n) -> Dict[str, Any]:
        if isinstance(n, oir.FieldAccess):
            off = n.offset
            if isinstance(off, common.CartesianOffset):
                o: Any = [int(off.i), int(off.j), int(off.k)]
            elif isinstance(off, oir.VariableKOffset):
                o = {"vk": self.expr(off.k)}
            elif isinstance(off, oir.AbsoluteKIndex):
                o = {"abs_k": int(off.k) if isinstance(off.k, int) else self.expr(off.k)}
            else:  # pragma: no cover
                raise NotImplementedError(f"offset {type(off)}")
            return {
                "t": "field",
                "name": str(n.name),
                "dtype": _dt(n.dtype),
                "off": o,
                "data_index": [self.expr(i) for i in n.data_index],
            }
        if isinstance(n, oir.ScalarAccess):
            return {"t": "scalar", "name": str(n.name), "dtype": _dt(n.dtype)}
        if isinstance(n, oir.Literal):
            v = n.value
            if isinstance(v, common.BuiltInLiteral):
                v = {"true": "True", "false": "False"}.get(v.value, v.value)
            return {"t": "lit", "value": str(v), "dtype": _dt(n.dtype)}
        if isinstance(n, oir.UnaryOp):
            return {"t": "unary", "op": str(n.op.value), "expr": self.expr(n.expr), "dtype": _dt(n.dtype)}
        if isinstance(n, oir.BinaryOp):
            return {
                "t": "binary",
                "op": str(n.op.value),
                "left": self.expr(n.left),
                "right": self.expr(n.right),
                "dtype": _dt(n.dtype),
            }
        if isinstance(n, oir.TernaryOp):
            return {
                "t": "ternary",
                "cond": self.expr(n.cond),
                "true": self.expr(n.true_expr),
                "false": self.expr(n.false_expr),
                "dtype": _dt(n.dtype),
            }
        if isinstance(n, oir.Cast):
            return {"t": "cast", "expr": self.expr(n.expr), "dtype": _dt(n.dtype)}
        if isinstance(n, oir.NativeFuncCall):
            return {
                "t": "call",
                "func": str(n.func.value),
                "args": [self.expr(a) for a in n.args],
                "dtype": _dt(n.dtype),
            }
        if isinstance(n, oir.IteratorAccess):
            return {"t": "iter", "axis": str(n.name.value), "dtype": _dt(n.dtype)}
        raise NotImplementedError(f"b200: OIR expression {type(n).__name__}")

    # -- statements -------------------------------------------------------------------------
    def stmt(self, n) -> Dict[str, Any]:
        if isinstance(n, oir.AssignStmt):
            return {"t": "assign", "left": self.expr(n.left), "right": self.expr(n.right)}
        if isinstance(n, oir.MaskStmt):
            return {"t": "mask", "mask": self.expr(n.mask), "body": [self.stmt(s) for s in n.body]}
        if isinstance(n, oir.While):
            return {"t": "while", "cond": self.expr(n.cond), "body": [self.stmt(s) for s in n.body]}
        if isinstance(n, oir.HorizontalRestriction):
            return {
                "t": "hregion",
                "i": [_bound(n.mask.i.start), _bound(n.mask.i.end)],
                "j": [_bound(n.mask.j.start), _bound(n.mask.j.end)],
                "body": [self.stmt(s) for s in n.body],
            }
        if isinstance(n, oir.CodeBlock):
            raise NotImplementedError("b200: oir.CodeBlock")
        raise NotImplementedError(f"b200: OIR statement {type(n).__name__}")

    def ext(self, e) -> list:
        return [[int(e[0][0]), int(e[0][1])], [int(e[1][0]), int(e[1][1])]]

    def run(self) -> Dict[str, Any]:
        st = self.stencil
        params = []
        for p in st.params:
            if isinstance(p, oir.FieldDecl):
                params.append(
                    {
                        "t": "field",
                        "name": str(p.name),
                        "dtype": _dt(p.dtype),
                        "dims": [bool(d) for d in p.dimensions],
                        "data_dims": [int(d) for d in p.data_dims],
                    }
                )
            else:
                params.append({"t": "scalar", "name": str(p.name), "dtype": _dt(p.dtype)})
        temps = []
        for d in st.declarations:
            e = self.field_extents.get(d.name)
            temps.append(
                {
                    "name": str(d.name),
                    "dtype": _dt(d.dtype),
                    "dims": [bool(x) for x in d.dimensions],
                    "data_dims": [int(x) for x in d.data_dims],
                    "extent": self.ext(e) if e is not None else [[0, 0], [0, 0]],
                }
            )
        loops = []
        for vl in st.vertical_loops:
            sections = []
            for sec in vl.sections:
                hes = []
                for he in sec.horizontal_executions:
                    hes.append(
                        {
                            "locals": [{"name": str(d.name), "dtype": _dt(d.dtype)} for d in he.declarations],
                            "extent": self.ext(self.block_extents[id(he)]),
                            "body": [self.stmt(s) for s in he.body],
                        }
                    )
                sections.append(
                    {"interval": [_bound(sec.interval.start), _bound(sec.interval.end)], "hes": hes}
                )
            caches = []
            for c in vl.caches:
                if isinstance(c, oir.KCache):
                    caches.append({"t": "k", "name": str(c.name), "fill": bool(c.fill), "flush": bool(c.flush)})
                else:
                    caches.append({"t": "ij", "name": str(c.name)})
            loops.append({"order": str(vl.loop_order.value), "sections": sections, "caches": caches})
        return {
            "t": "stencil",
            "ir_version": b2ir.IR_VERSION,
            "name": str(st.name),
            "params": params,
            "temporaries": temps,
            "loops": loops,
        }


def lower_oir(stencil: oir.Stencil) -> Dict[str, Any]:
    """Lower a (pipeline-processed) `oir.Stencil` to the b200 IR, annotated with extents."""
    return _Lower(stencil).run()


def args_data_to_ir(args_data) -> Dict[str, Any]:
    """Serialise gt4py's ModuleData (reference: backend/module_generator.py:31-106)."""
    finfo: Dict[str, Any] = {}
    for name, fi in args_data.field_info.items():
        finfo[name] = {
            "access": fi.access.name,
            "boundary": [[int(lo), int(hi)] for lo, hi in fi.boundary],
            "axes": list(fi.axes),
            "data_dims": [int(d) for d in fi.data_dims],
            "dtype": str(fi.dtype),
        }
    pinfo: Dict[str, Any] = {}
    for name, pi in args_data.parameter_info.items():
        pinfo[name] = {"access": pi.access.name, "dtype": str(pi.dtype)}
    return {
        "field_info": finfo,
        "parameter_info": pinfo,
        "domain_info": {"min_k": int(args_data.domain_info.min_sequential_axis_size)},
    }


def default_pipeline(variant: str = "default"):
    """OIR pipelines the b200 backend uses.

    "default": the reference's full DefaultPipeline (what `gt:gpu` runs, backend/gtcpp_backend.py:41-47).
    "staged" : the same without OnTheFlyMerging, so multi-stage PARALLEL blocks keep their
               temporaries (with IJ extents) and the emitter can tile them on-chip instead of
               recomputing them per point.
    """
    from gt4py.cartesian.gtc.passes.oir_optimizations.horizontal_execution_merging import OnTheFlyMerging
    from gt4py.cartesian.gtc.passes.oir_pipeline import DefaultPipeline

    if variant == "default":
        return DefaultPipeline()
    if variant == "staged":
        return DefaultPipeline(skip=[OnTheFlyMerging])
    raise ValueError(variant)


def lower_definition(definition, *, name=None, externals=None, dtypes=None, variant="default", **build_opts):
    """Frontend + passes + lowering for a GTScript definition function (dev/fixture helper)."""
    from gt4py.cartesian import backend as gt_backend
    from gt4py.cartesian.backend.module_generator import make_args_data_from_gtir
    from gt4py.cartesian.definitions import BuildOptions
    from gt4py.cartesian.gtc.gtir_to_oir import GTIRToOIR
    from gt4py.cartesian.stencil_builder import StencilBuilder
    from gt4py.cartesian import gtscript

    if dtypes:
        gtscript._set_arg_dtypes(definition, dtypes)
    opts = BuildOptions(name=name or definition.__name__, module=definition.__module__, **build_opts)
    builder = StencilBuilder(definition, backend=gt_backend.from_name("numpy"), options=opts)
    if externals:
        builder = builder.with_externals(externals)
    base_oir = GTIRToOIR().visit(builder.gtir)
    out = lower_oir(default_pipeline(variant).run(base_oir))
    out.update(args_data_to_ir(make_args_data_from_gtir(builder.gtir_pipeline)))
    out["variant"] = variant
    out["options"] = {k: v for k, v in build_opts.items() if isinstance(v, (int, float, str, bool))}
    return out
