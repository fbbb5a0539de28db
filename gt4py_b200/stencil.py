"""Stand-alone stencil object of the b200 backend (host-side mirror of gt4py's `StencilObject`).

Where gt4py is installed, `backend="b200"` produces a real `gt4py.cartesian.StencilObject` subclass
(see `backend.py`).  This module provides the same call surface *without* gt4py, built from a
serialised stencil IR, so that a machine that only has the lowered IR (tests/golden/ir/*.json), the
launcher and a GPU can run, test and benchmark the hot path.  It restates the run-path logic of the
reference:

* `__call__(fields…, params…, domain=None, origin=None, validate_args=True, exec_info=None)`
                                         (reference: backend/templates/stencil_module.py.in:91-158)
* `_call_run`: array extraction incl. `__gt_dims__` / `__gt_origin__`, origin normalisation, maximum
  domain inference, validation with the reference's error types and messages, per-call cache
                                         (reference: stencil_object.py:69-93, 296-612)
* `freeze(origin=…, domain=…)` -> `FrozenStencil` skipping all of the above
                                         (reference: stencil_object.py:96-136, 614-643)
"""

from __future__ import annotations

import sys
import time
import warnings
from typing import Any, Dict, Optional, Sequence, Tuple

import numpy as np

from . import runtime, storage as b2storage


class FieldInfo:
    def __init__(self, d: Dict[str, Any]):
        self.access = d["access"]
        self.boundary = tuple((int(lo), int(hi)) for lo, hi in d["boundary"])
        self.axes = tuple(d["axes"])
        self.data_dims = tuple(int(x) for x in d["data_dims"])
        self.dtype = np.dtype(d["dtype"])
        self.domain_mask = tuple(a in self.axes for a in "IJK")
        self.domain_ndim = len(self.axes)
        self.ndim = len(self.axes) + len(self.data_dims)

    def __repr__(self):
        return f"FieldInfo(access={self.access}, boundary={self.boundary}, axes={self.axes}, data_dims={self.data_dims}, dtype={self.dtype})"


class ParameterInfo:
    def __init__(self, d: Dict[str, Any]):
        self.access = d["access"]
        self.dtype = np.dtype(d["dtype"])


def _filter_mask(seq, mask):
    return tuple(s for s, m in zip(seq, mask) if m)


class _ArgInfo:
    __slots__ = ("view", "origin", "dimensions")

    def __init__(self, view, origin, dimensions):
        self.view, self.origin, self.dimensions = view, origin, dimensions


def extract_array_infos(field_args: Dict[str, Any]) -> Dict[str, Optional[_ArgInfo]]:
    """reference: stencil_object.py:69-93 (without cupy): honour __gt_dims__ / __gt_origin__."""
    out: Dict[str, Optional[_ArgInfo]] = {}
    for name, arg in field_args.items():
        if arg is None:
            out[name] = None
            continue
        view = runtime.as_view(arg)
        dims = getattr(arg, "__gt_dims__", None)
        if dims is not None:
            dims = tuple(str(d) for d in dims)
            sorted_dims = [d for d in "IJK" if d in dims]
            sorted_dims += [str(d) for d in sorted(int(d) for d in dims if str(d).isdigit())]
            view = view.transpose([dims.index(d) for d in sorted_dims])
            dims = tuple(sorted_dims)
        origin = getattr(arg, "__gt_origin__", None)
        out[name] = _ArgInfo(view, tuple(int(o) for o in origin) if origin is not None else None, dims)
    return out


class B200Stencil:
    """A compiled stencil callable like a gt4py StencilObject."""

    backend = "b200"

    def __init__(self, stencil_ir: Dict[str, Any], options: Optional[Dict[str, Any]] = None, *, name: Optional[str] = None):
        opts = dict(options or {})
        self._init_from_compiled(runtime.CompiledStencil(stencil_ir, opts, name=name or stencil_ir["name"]), opts)

    def _init_from_compiled(self, compiled: "runtime.CompiledStencil", options: Dict[str, Any]) -> None:
        """(also the entry for the gt4py plug-in: a mirror of a `backend="b200"` StencilObject around ITS compiled stencil)"""
        stencil_ir = compiled.ir
        self.ir = stencil_ir
        self.name = compiled.name
        self.backend_options = dict(options)
        self.compiled = compiled
        self.field_info: Dict[str, Optional[FieldInfo]] = {
            n: (FieldInfo(fi) if fi is not None else None) for n, fi in stencil_ir["field_info"].items()
        }
        self.parameter_info: Dict[str, Optional[ParameterInfo]] = {
            n: (ParameterInfo(pi) if pi is not None else None) for n, pi in stencil_ir["parameter_info"].items()
        }
        self.min_k = int(stencil_ir["domain_info"]["min_k"])
        self._signature = [p["name"] for p in stencil_ir["params"]]
        self._field_names = [p["name"] for p in stencil_ir["params"] if p["t"] == "field"]
        self._param_names = [p["name"] for p in stencil_ir["params"] if p["t"] == "scalar"]
        # arguments pruned by the frontend still belong to the call signature
        for n in stencil_ir["field_info"]:
            if n not in self._signature:
                self._signature.append(n)
                self._field_names.append(n)
        for n in stencil_ir["parameter_info"]:
            if n not in self._signature:
                self._signature.append(n)
                self._param_names.append(n)
        self._cache: Dict[int, Tuple[tuple, dict]] = {}
        self.device_sync = bool(self.backend_options.get("device_sync", True))

    # ---- normalisation helpers (reference: stencil_object.py:265-340, 497-529) --------------------
    @staticmethod
    def _make_origin_dict(origin) -> Dict[str, Tuple[int, ...]]:
        try:
            if isinstance(origin, dict):
                return dict(origin)
            if origin is None:
                return {}
            if isinstance(origin, (tuple, list)) and all(isinstance(o, (int, np.integer)) for o in origin):
                return {"_all_": tuple(int(o) for o in origin)}
        except Exception:
            pass
        raise ValueError(f"Invalid 'origin' value ({origin})")

    def _normalize_origins(self, infos, origin) -> Dict[str, Tuple[int, ...]]:
        origin = self._make_origin_dict(origin)
        all_origin = origin.get("_all_")
        for name, fi in self.field_info.items():
            if fi is None:
                continue
            assert name in infos, f"Missing value for '{name}' field."
            fo = origin.get(name)
            if fo is not None:
                if len(fo) != fi.ndim:
                    assert len(fo) == fi.domain_ndim, f"Invalid origin specification ({fo}) for '{name}' field."
                    origin[name] = (*fo, *((0,) * len(fi.data_dims)))
            elif all_origin is not None:
                origin[name] = (*_filter_mask(all_origin, fi.domain_mask), *((0,) * len(fi.data_dims)))
            elif infos.get(name) is not None and infos[name].origin is not None:
                origin[name] = infos[name].origin
            else:
                origin[name] = (0,) * fi.ndim
        return origin

    def _get_max_domain(self, infos, origin, *, only=None, squeeze=True) -> Tuple[int, ...]:
        big = sys.maxsize
        max_domain = [big, big, big]
        for name, fi in self.field_info.items():
            if fi is None or fi.access == "NONE" or (only is not None and name != only):
                continue
            info = infos.get(name)
            assert info is not None, f"Invalid value for '{name}' field."
            upper = [b[1] for b, m in zip(fi.boundary, fi.domain_mask) if m]
            fo = origin[name]
            ax = 0
            for a in range(3):
                if fi.domain_mask[a]:
                    max_domain[a] = min(max_domain[a], info.view.shape[ax] - (fo[ax] + upper[ax]))
                    ax += 1
        if squeeze:
            return tuple(d if d != big else 1 for d in max_domain)
        return tuple(max_domain)

    def _validate_args(self, infos, params, domain, origin) -> None:
        """reference: stencil_object.py:342-494 — same checks, same exception types/messages."""
        if len(domain) != 3:
            raise ValueError(f"Invalid 'domain' value '{domain}'")
        try:
            domain = tuple(int(d) for d in domain)
        except Exception as ex:
            raise ValueError(f"Invalid 'domain' value ({domain})") from ex
        if not all(d > 0 for d in domain):
            raise ValueError(f"Compute domain contains zero sizes '{domain}')")
        max_domain = self._get_max_domain(infos, origin, squeeze=False)
        if not all(d <= m for d, m in zip(domain, max_domain)):
            offending = []
            for name, fi in self.field_info.items():
                if fi is None or fi.access == "NONE":
                    continue
                used = self._get_max_domain(infos, origin, only=name, squeeze=False)
                if any(u < d for u, d in zip(used, domain)):
                    offending.append((name, used))
            raise ValueError(
                f"Compute domain too large for stencil {self.name}: \n"
                f"  Stencil domain is {domain} but field indexation leads to read outside of bounds.\n"
                f"  Check region/horizontal offsets or interval/vertical offsets, or stencil domain.\n"
                f"  Offending fields (name, size with offset removed): {offending}"
            )
        if domain[2] < self.min_k:
            raise ValueError(
                f"Compute domain too small. Sequential axis is {domain[2]}, but must be at least {self.min_k}."
            )
        for name, fi in self.field_info.items():
            if fi is None or fi.access == "NONE":
                continue
            if name not in infos or infos[name] is None:
                raise ValueError(f"Missing value for '{name}' field.")
            info = infos[name]
            dims = tuple(list(fi.axes) + [str(d) for d in range(len(fi.data_dims))])
            if not _strides_follow_layout(info.view.strides, dims):
                warnings.warn(
                    f"The layout of the field '{name}' is not recommended for this backend."
                    f"This may lead to performance degradation. Please consider using the"
                    f"provided allocators in `gt4py.storage`.",
                    stacklevel=3,
                )
            if info.view.dtype != fi.dtype:
                raise TypeError(f"The dtype of field '{name}' is '{info.view.dtype}' instead of '{fi.dtype}'")
            if info.view.ndim != fi.domain_ndim + len(fi.data_dims):
                raise ValueError(
                    f"Storage for '{name}' has {info.view.ndim} dimensions but the API signature "
                    f"expects {fi.domain_ndim + len(fi.data_dims)} ('{fi.axes}[{fi.data_dims}]')"
                )
            if info.dimensions is not None and dims != info.dimensions:
                raise ValueError(
                    f"Storage for '{name}' has dimensions '{info.dimensions}' but the API signature "
                    f"expects '[{', '.join(fi.axes)}]'"
                    + (f" and {len(fi.data_dims)}" if fi.data_dims else "")
                )
            if tuple(info.view.shape[fi.domain_ndim :]) != fi.data_dims:
                raise ValueError(
                    f"Field '{name}' expects data dimensions {fi.data_dims} but got {info.view.shape[fi.domain_ndim:]}"
                )
            lower = [b[0] for b, m in zip(fi.boundary, fi.domain_mask) if m]
            upper = [b[1] for b, m in zip(fi.boundary, fi.domain_mask) if m]
            fo = origin[name][: fi.domain_ndim]
            if any(o < lo for o, lo in zip(fo, lower)):
                full_min = tuple(b[0] if m else 0 for b, m in zip(fi.boundary, fi.domain_mask))
                raise ValueError(f"Origin for field {name} too small. Must be at least {full_min}, is {tuple(fo)}")
            spatial = _filter_mask(domain, fi.domain_mask)
            min_shape = tuple(lb + d + ub for lb, d, ub in zip(lower, spatial, upper))
            if min_shape > tuple(info.view.shape):
                raise ValueError(
                    f"Shape of field {name} is {info.view.shape} but must be at least {min_shape} for given domain and origin."
                )
        for name, pi in self.parameter_info.items():
            if pi is None or pi.access == "NONE":
                continue
            if name not in params:
                raise ValueError(f"Missing value for '{name}' parameter.")
            if np.dtype(type(params[name])) != pi.dtype:
                raise TypeError(f"The type of parameter '{name}' is '{type(params[name])}' instead of '{pi.dtype}'")

    # ---- call path -------------------------------------------------------------------------------
    def __call__(self, *args, domain=None, origin=None, validate_args=True, exec_info=None, **kwargs):
        if exec_info is not None:
            exec_info["call_start_time"] = time.perf_counter()
        if len(args) > len(self._signature):
            raise TypeError(f"{self.name}() takes {len(self._signature)} positional arguments but {len(args)} were given")
        bound = dict(zip(self._signature, args))
        for k, v in kwargs.items():
            if k in bound:
                raise TypeError(f"{self.name}() got multiple values for argument '{k}'")
            if k not in self._signature:
                raise TypeError(f"{self.name}() got an unexpected keyword argument '{k}'")
            bound[k] = v
        field_args = {n: bound.get(n) for n in self._field_names}
        param_args = {n: bound[n] for n in self._param_names if n in bound}
        self._call_run(field_args, param_args, domain, origin, validate_args=validate_args, exec_info=exec_info)
        if exec_info is not None:
            exec_info["call_end_time"] = time.perf_counter()

    def _call_run(self, field_args, parameter_args, domain, origin, *, validate_args=True, exec_info=None):
        if exec_info is not None:
            exec_info["call_run_start_time"] = time.perf_counter()
        from . import hostpipe

        if any(a is not None and hostpipe.is_host_array(a) for a in field_args.values()):
            # arguments in HOST memory: staged through device storages (H2D, kernels, D2H; K-slab pipelined when possible)
            hostpipe.host_call(self, field_args, parameter_args, domain, origin, validate_args=validate_args, exec_info=exec_info)
            if exec_info is not None:
                exec_info["call_run_end_time"] = time.perf_counter()
            return
        infos = extract_array_infos(field_args)
        key = hash(
            (
                tuple((n, i.view.shape, i.origin or (0, 0, 0)) for n, i in infos.items() if i is not None),
                *parameter_args.keys(),
                repr(domain),
                repr(origin),
            )
        )
        cached = self._cache.get(key)
        if cached is None:
            origin = self._normalize_origins(infos, origin)
            if domain is None:
                domain = self._get_max_domain(infos, origin)
            if validate_args:
                self._validate_args(infos, parameter_args, domain, origin)
            self._cache[key] = (tuple(domain), origin)
        else:
            domain, origin = cached
        views = {n: (i.view if i is not None else None) for n, i in infos.items()}
        self.run(_domain_=tuple(domain), _origin_=origin, exec_info=exec_info, **views, **parameter_args)
        if exec_info is not None:
            exec_info["call_run_end_time"] = time.perf_counter()

    def run(self, _domain_, _origin_, exec_info=None, *, stream=None, subbox=None, **args):
        """reference: stencil_module.py.in:160-169 (native call + optional device sync)."""
        if exec_info is not None:
            exec_info["domain"] = _domain_
            exec_info["origin"] = _origin_
            exec_info["run_start_time"] = time.perf_counter()
        fields = {n: args.get(n) for n in self._field_names}
        params = {n: args[n] for n in self._param_names if n in args}
        timer = None
        if exec_info is not None:  # device-side duration of the call (see runtime.DeviceTimer)
            timer, sh = runtime.DeviceTimer(), (stream if stream is not None else runtime.current_stream_handle())
            timer.start(sh)
        n = self.compiled.run(fields, params, _domain_, _origin_, stream=stream, subbox=subbox)
        if timer is not None:
            timer.stop(sh)
        if self.device_sync:
            runtime.check(runtime.load_library().b200_stream_synchronize(stream if stream is not None else runtime.current_stream_handle()))
        if exec_info is not None:
            exec_info["run_end_time"] = time.perf_counter()
            exec_info["b200_kernel_launches"] = n
            if self.device_sync:
                exec_info["run_device_time"] = timer.elapsed()
            else:
                exec_info["b200_device_timer"] = timer
        return n

    # ---- autotuning (tile shapes are a property of stencil x domain x device) --------
    #: code-generation variants tried by autotune(); "auto" pitch = the common row pitch of the arguments
    DEFAULT_CANDIDATES = (
        {},
        {"static_pitch": "auto"},
        {"interior_loop": True},
        {"interior_loop": True, "static_pitch": "auto"},
        {"interior_loop": True, "static_pitch": "auto", "l2_prefetch": 4},
        {"interior_loop": True, "static_pitch": "auto", "l2_prefetch": 1},
        {"interior_loop": True, "static_pitch": "auto", "tile_j": 128},
        {"interior_loop": True, "static_pitch": "auto", "tile_j": 32},
        {"interior_loop": True, "static_pitch": "auto", "warps": 8},
        {"interior_loop": True, "static_pitch": "auto", "warps": 2},
        {"interior_loop": True, "static_pitch": "auto", "prefetch": 0},
        {"interior_loop": True, "static_pitch": "auto", "vector_width": 4},
        {"interior_loop": True, "static_pitch": "auto", "vector_width": 4, "prefetch": 0},
        {"interior_loop": "steady", "static_pitch": "auto"},
        {"interior_loop": True, "static_pitch": "auto", "row_pointers": True},
        {"interior_loop": "steady", "static_pitch": "auto", "row_pointers": True},
        {"static_pitch": "auto", "row_pointers": True},
        {"interior_loop": True, "static_pitch": "auto", "stcs": True},
        {"interior_loop": True, "static_pitch": "auto", "stcs": True, "ldcs": True},
        {"interior_loop": True, "static_pitch": "auto", "tile_j": 52},
        {"interior_loop": True, "static_pitch": "auto", "tile_j": 96},
        {"interior_loop": True, "static_pitch": "auto", "min_blocks": 8},
        {"interior_loop": True, "static_pitch": "auto", "min_blocks": 8, "tile_j": 128},
        {"interior_loop": True, "static_pitch": "auto", "min_blocks": 9},
        {"interior_loop": True, "static_pitch": "auto", "tile_j": 256},
        # bulk-async variants (codegen_stream.py: bulk / tensor-map copies into a per-warp shared-memory ring, LDS reads)
        {"interior_loop": True, "static_pitch": "auto", "tma": 3, "tile_j": 32, "prefetch": 1, "tma_mode": "bulk"},
        {"interior_loop": True, "static_pitch": "auto", "tma": 3, "tile_j": 32, "prefetch": 1},
        {"interior_loop": True, "static_pitch": "auto", "tma": 4, "tile_j": 32, "prefetch": 1, "tma_mode": "bulk"},
        {"interior_loop": True, "static_pitch": "auto", "tma": 3, "tile_j": 48, "prefetch": 1, "tma_mode": "bulk"},
        {"interior_loop": True, "static_pitch": "auto", "tma": 3, "tile_j": 24, "prefetch": 1, "tma_mode": "bulk"},
        {"interior_loop": True, "static_pitch": "auto", "tma": 3, "tile_j": 16, "prefetch": 1},
        {"interior_loop": True, "static_pitch": "auto", "tma": 4, "prefetch": 1, "tma_mode": "bulk"},
        {"interior_loop": True, "static_pitch": "auto", "tma": 3, "tile_j": 32, "prefetch": 1, "l2_prefetch": 4},
        {"interior_loop": True, "static_pitch": "auto", "tma": 3, "tile_j": 32, "prefetch": 1, "tma_mode": "bulk", "stcs": True},
        {"interior_loop": True, "static_pitch": "auto", "tma": 3, "tile_j": 32, "prefetch": 0},
        {"interior_loop": True, "static_pitch": "auto", "tma": 3, "tile_j": 32, "prefetch": 0, "tma_mode": "bulk"},
        {"interior_loop": True, "static_pitch": "auto", "tma": 3, "tile_j": 32, "prefetch": 1, "tma_mode": "bulk", "tma_rows": 8, "tma_smem_kb": 48},
        {"interior_loop": True, "static_pitch": "auto", "uniform_task": False},  # (the round-1 addressing, for the record)
        {"interior_loop": True, "static_pitch": "auto", "min_blocks": 10},
        {"interior_loop": True, "static_pitch": "auto", "min_blocks": 12},
        {"interior_loop": True, "static_pitch": "auto", "prefetch": 2},
        {"edge_loop": True},
        {"pure_loop": False},
        {"vector_width": 4},
        {"tile_j": 32},
        {"tile_j": 128},
        {"l2_prefetch": 4},
        {"l2_prefetch": 1},
        {"prefetch": 0},
        {"warps": 2},
        {"warps": 8},
    )

    @staticmethod
    def resolve_candidate(cand: Dict[str, Any], views: Dict[str, Any]) -> Optional[Dict[str, Any]]:
        """Fill argument-dependent values of a candidate: static_pitch="auto" -> the J stride shared
        by all 3-D arguments (None when they do not share one)."""
        cand = dict(cand)
        if cand.get("static_pitch") == "auto":
            pitches = {v.strides[1] for v in views.values() if v is not None and v.ndim >= 3 and v.strides[0] == 1}
            if len(pitches) != 1:
                return None
            cand["static_pitch"] = int(pitches.pop())
        return cand

    def autotune(self, fields, params, *, domain, origin, candidates=None, iters: int = 10, verbose: bool = False, refine: int = 3):
        """Time the code-generation variants in `candidates` (option dicts merged over the current
        options) on the given device arguments and keep the fastest one whose written fields are
        bit-identical to the first candidate's (the current options).  Outputs are overwritten.
        Returns [(options, ms_per_launch)] sorted by time."""
        import torch

        base = dict(self.backend_options)
        views = {n: (runtime.as_view(fields[n]) if fields.get(n) is not None else None) for n in self._field_names}
        written = [n for n, fi in self.field_info.items() if fi is not None and fi.access in ("WRITE", "READ_WRITE") and views.get(n) is not None]
        # stencils that update a field in place (the Thomas solver's sup / rhs, the w of an implicit solve): every candidate
        # starts from a saved copy of those fields, is validated after ONE application, and the caller's values are put
        # back at the end; the timed launches run on whatever the repeated application leaves (the generated kernels have no
        # data-dependent control flow apart from the guarded division fallbacks)
        inplace = [n for n in written if self.field_info[n].access == "READ_WRITE"]

        def tensor_of(n):
            t = fields[n].torch() if isinstance(fields[n], b2storage.DeviceArray) else fields[n]
            return t if hasattr(t, "clone") and hasattr(t, "zero_") else None

        if inplace and any(tensor_of(n) is None for n in inplace):
            raise ValueError("autotune: in-place fields must be device storages or tensors (they are saved and restored)")
        saved = {n: tensor_of(n).clone() for n in inplace}

        def restore():
            for n, t in saved.items():
                tensor_of(n).copy_(t)

        def tensors():
            ts = [tensor_of(n) for n in written]
            return [t for t in ts if t is not None]

        def snapshot():
            return [t.clone() for t in tensors()]

        results, rejected = [], []
        seen = set()
        expect = None
        for cand in candidates if candidates is not None else self.DEFAULT_CANDIDATES:
            cand = self.resolve_candidate(cand, views)
            if cand is None:
                continue
            opts = {**base, **cand}
            try:
                cs = runtime.CompiledStencil(self.ir, opts, name=self.name)
                cs._specialize = "off"  # candidates are explicit variants: no lazy re-specialisation on top of them
            except Exception:  # a variant that does not apply to this stencil
                continue
            if cs.source in seen:
                continue
            seen.add(cs.source)
            scal = cs.pack_scalars(params)
            descs = cs.make_field_descs(views, origin)
            for t in tensors():
                t.zero_()  # a variant that writes nothing must not inherit the previous one's result
            restore()
            for _ in range(1 if inplace else 3):
                cs.run_descs(descs, scal, domain)
            got = snapshot()
            if expect is None:
                expect = got  # the current options are the reference: parity-tested against the oracle
            elif any(not torch.equal(a, b) for a, b in zip(expect, got)):
                rejected.append(cand)
                continue
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                cs.run_descs(descs, scal, domain)
            e1.record()
            e1.synchronize()
            results.append((cand, e0.elapsed_time(e1) / iters, cs, opts))
            if verbose:
                print(f"autotune {self.name}: {cand} -> {results[-1][1]:.4f} ms")
        if not results:
            raise RuntimeError("autotune: no candidate could be built")
        results.sort(key=lambda r: r[1])
        # the sweep's 10-launch timings are a few per cent noisy: re-time the leaders (interleaved, best of
        # `refine` rounds) so that the choice among near-equal variants is not a coin toss
        if refine and len(results) > 1:
            top = results[: min(6, len(results))]
            best_ms = {id(r[2]): float("inf") for r in top}  # only the interleaved re-timings rank the leaders
            for _ in range(int(refine)):
                for cand, _ms, cs, _opts in top:
                    scal, descs = cs.pack_scalars(params), cs.make_field_descs(views, origin)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _i in range(2 * iters):
                        cs.run_descs(descs, scal, domain)
                    e1.record()
                    e1.synchronize()
                    best_ms[id(cs)] = min(best_ms[id(cs)], e0.elapsed_time(e1) / (2 * iters))
            results = sorted([(r[0], best_ms[id(r[2])], r[2], r[3]) for r in top], key=lambda r: r[1]) + results[len(top):]
        if inplace:
            torch.cuda.synchronize()
            restore()
        best = results[0]
        self.compiled, self.backend_options = best[2], best[3]
        self.tuned = [(r[0], round(r[1], 5)) for r in results]
        self.tune_rejected = rejected
        return self.tuned

    def autotune_isolated(self, fields, params, *, domain, origin, candidates=None, iters: int = 10, timeout: float = 300.0,
                          device: Optional[int] = None):  # fmt: skip
        """`autotune` with the candidate sweep in a sacrificial child process (gt4py_b200/tune_worker.py) on
        synthetic arguments of the same geometry: a variant that faults or hangs on the device cannot take
        the caller's CUDA context with it.  The parent then walks the child's ranking and adopts the first
        variant that is bit-identical to the current one on the REAL arguments (outputs are overwritten).
        Raises if the child fails; the current options stay in force in that case."""
        import json
        import os
        import subprocess
        import tempfile

        import torch

        views = {n: (runtime.as_view(fields[n]) if fields.get(n) is not None else None) for n in self._field_names}
        spec = {
            "name": self.name, "ir": self.ir, "options": self.backend_options, "params": {k: (v.item() if hasattr(v, "item") else v) for k, v in params.items()},
            "domain": [int(d) for d in domain], "origin": {k: [int(x) for x in v] for k, v in origin.items()},
            "candidates": [dict(c) for c in (candidates if candidates is not None else self.DEFAULT_CANDIDATES)],
            "iters": int(iters), "device": int(torch.cuda.current_device() if device is None else device),
            "fields": {n: (None if v is None else {"shape": list(v.shape), "strides": list(v.strides), "dtype": v.dtype.name, "phase": v.ptr % 256})
                       for n, v in views.items()},
        }  # fmt: skip
        with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as fh:
            json.dump(spec, fh)
            spec_path = fh.name
        try:
            env = dict(os.environ)
            root = str(__import__("pathlib").Path(__file__).resolve().parent.parent)
            env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
            for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT", "TORCHELASTIC_RUN_ID"):
                env.pop(k, None)  # the child is a plain single-device process
            proc = subprocess.run([sys.executable, "-m", "gt4py_b200.tune_worker", spec_path], capture_output=True, text=True,
                                  timeout=timeout, env=env)  # fmt: skip
        finally:
            try:
                os.unlink(spec_path)
            except OSError:
                pass
        if proc.returncode != 0:
            raise RuntimeError(f"autotune worker exited with {proc.returncode}: {proc.stderr[-600:]}")
        res = json.loads(proc.stdout.strip().splitlines()[-1])
        ranking = [(c, float(ms)) for c, ms in res["tuned"]]
        self.tune_rejected = list(res.get("rejected", []))
        written = [n for n, fi in self.field_info.items() if fi is not None and fi.access in ("WRITE", "READ_WRITE") and views.get(n) is not None]

        def outputs():
            ts = [fields[n].torch() if isinstance(fields[n], b2storage.DeviceArray) else fields[n] for n in written]
            return [t for t in ts if hasattr(t, "clone") and hasattr(t, "zero_")]

        def run_and_snapshot(cs):
            for t in outputs():
                t.zero_()
            cs.run_descs(cs.make_field_descs(views, origin), cs.pack_scalars(params), domain)
            torch.cuda.synchronize()
            return [t.clone() for t in outputs()]

        expect = run_and_snapshot(self.compiled)
        base = dict(self.backend_options)
        for cand, _ms in ranking:
            opts = {**base, **cand}
            cs = runtime.CompiledStencil(self.ir, opts, name=self.name)  # cubin: disk cache written by the child
            got = run_and_snapshot(cs)
            if all(torch.equal(a, b) for a, b in zip(expect, got)):
                self.compiled, self.backend_options = cs, opts
                break
            self.tune_rejected.append({"rejected_on_real_arguments": cand})
        self.tuned = [(c, round(ms, 5)) for c, ms in ranking]
        return self.tuned

    def freeze(self, *, origin: Dict[str, Tuple[int, ...]], domain: Tuple[int, ...]) -> "FrozenStencil":
        return FrozenStencil(self, origin, tuple(domain))

    def clean_call_args_cache(self) -> None:
        self._cache.clear()


class FrozenStencil:
    """Pre-resolved origin/domain and pre-built field descriptors: the low-overhead launch path
    (reference: stencil_object.py:96-136).  Descriptors are rebuilt only when an argument's
    (pointer, shape, strides) changes."""

    def __init__(self, stencil: B200Stencil, origin, domain):
        for name, fi in stencil.field_info.items():
            if fi is None:
                continue
            if name not in origin or len(origin[name]) != fi.ndim:
                raise ValueError(f"'{name}' origin {origin.get(name)} is not a {fi.ndim}-dimensional integer tuple")
        self.stencil_object = stencil
        self.origin = origin
        self.domain = domain
        # argument objects -> prepared descriptors.  Keyed by object identity (the entry holds the objects, so an id
        # cannot be recycled while it is cached) and re-validated against the arguments' pointers: a model step
        # calls with the same few buffer sets over and over, and this path is what it pays per launch.
        self._cache: Dict[tuple, tuple] = {}
        self._scal_key = None
        self._scal = None
        self._names = tuple(stencil._field_names)
        self._pnames = tuple(stencil._param_names)

    def _descs_for(self, kwargs):
        st = self.stencil_object
        args = tuple(kwargs.get(n) for n in self._names)
        ids = tuple(map(id, args))
        hit = self._cache.get(ids)
        if hit is not None:
            descs, held, ptrs = hit
            # same objects: geometry can only have changed if a tensor was re-pointed in place
            if all(a is h for a, h in zip(args, held)) and ptrs == tuple((_ptr_of(a), getattr(a, "shape", None)) for a in args):
                return descs
        views = {n: (runtime.as_view(a) if a is not None else None) for n, a in zip(self._names, args)}
        descs = st.compiled.make_field_descs(views, self.origin)
        if len(self._cache) >= 16:
            self._cache.clear()
        self._cache[ids] = (descs, args, tuple((_ptr_of(a), getattr(a, "shape", None)) for a in args))
        return descs

    def __call__(self, *, exec_info=None, stream=None, subbox=None, halo_wait=None, **kwargs) -> int:
        st = self.stencil_object
        descs = self._descs_for(kwargs)
        if self._pnames:
            skey = tuple(kwargs.get(n) for n in self._pnames)
            if skey != self._scal_key:
                self._scal = st.compiled.pack_scalars({n: kwargs.get(n) for n in self._pnames})
                self._scal_key = skey
        elif self._scal is None:
            self._scal = st.compiled.pack_scalars({})
        n = st.compiled.run_descs(descs, self._scal, self.domain, stream=stream, subbox=subbox, halo_wait=halo_wait)
        if st.device_sync:
            runtime.check(runtime.load_library().b200_stream_synchronize(stream if stream is not None else runtime.current_stream_handle()))
        return n


def _ptr_of(obj) -> int:
    """cheap identity of the memory an argument points at (no view construction)"""
    if obj is None:
        return 0
    dp = getattr(obj, "data_ptr", None)
    if dp is not None:
        return dp() if callable(dp) else dp
    cai = getattr(obj, "__cuda_array_interface__", None)
    return int(cai["data"][0]) if cai is not None else id(obj)


def _strides_follow_layout(strides, dims) -> bool:
    lm = b2storage.layout_map(dims)
    if len(strides) != len(lm):
        return False
    stride = 0
    for dim in reversed(np.argsort(lm)):
        if strides[dim] < stride:
            return False
        stride = strides[dim]
    return True


def from_ir_file(path, options=None) -> B200Stencil:
    from . import ir as b2ir

    return B200Stencil(b2ir.load_file(path), options)
