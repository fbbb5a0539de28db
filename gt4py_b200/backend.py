"""gt4py plug-in: registers `backend="b200"` through gt4py's own backend and storage registries.

Importing this module (done by `import gt4py_b200` whenever gt4py is importable) makes

    @gtscript.stencil(backend="b200")            /  gt4py.storage.zeros(..., backend="b200")

work unchanged for existing GTScript code.  What is replaced, and only this:

* the backend code generator / builder       `B200Backend.generate()`
      frontend + GTIR/OIR passes are gt4py's own (reference: stencil_builder.py:253-266,
      gtc/gtir_to_oir.py, gtc/passes/oir_pipeline.py); the OIR is lowered to the b200 IR
      (`from_oir.py`), turned into sm_100a CUDA by `codegen*.py`, compiled to a cubin by `jit.py`
      (instead of GridTools C++ + pybind11 + setuptools: backend/gtcpp_backend.py:35-62,
      backend/pyext_builder.py)
* the StencilObject run / launch path        `B200StencilObject`, generated `run()`
      cupy-free array extraction (reference: stencil_object.py:69-93 asserts cupy for GPU devices),
      native call through the C-ABI launcher (reference: backend/gtc_common.py:144-168)
* the gt4py.storage device allocator         `storage.py` hooked into storage/cartesian/utils.py
      (reference: `allocate_gpu`, `asarray`, `cpu_copy`, storage/cartesian/utils.py:168-279)

Everything else — caching (stencil id, cache-info validation), the module template, argument
validation, exec_info bookkeeping — is the reference's, reused as a library.
"""

from __future__ import annotations

import ctypes
import hashlib
import json
import pathlib
import threading
import time
from typing import Any, ClassVar, Dict, Optional

import numpy as np

from gt4py import storage as gt_storage
from gt4py.cartesian import backend as gt_backend
from gt4py.cartesian.backend.base import BaseBackend
from gt4py.cartesian.backend.module_generator import BaseModuleGenerator, make_args_data_from_gtir
from gt4py.cartesian.gtc import passes
from gt4py.cartesian.gtc.gtir_to_oir import GTIRToOIR
from gt4py.cartesian.stencil_object import ArgsInfo, StencilObject, _compute_domain_origin_cache_key
from gt4py.storage.cartesian import layout as gt_layout, utils as gt_storage_utils

from . import from_oir, hostpipe, ir as b2ir, runtime, storage as b2storage

# ---- storage hooks -----------------------------------------------------------------------------
_orig_allocate_gpu = gt_storage_utils.allocate_gpu
_orig_asarray = gt_storage_utils.asarray
_orig_cpu_copy = gt_storage_utils.cpu_copy


def _allocate_gpu(shape, layout_map, dtype, alignment_bytes, aligned_index):
    """Drop-in for storage/cartesian/utils.py:_allocate_gpu without cupy: torch-backed pitched buffer."""
    if gt_storage_utils.cp is not None:
        return _orig_allocate_gpu(shape, layout_map, dtype, alignment_bytes, aligned_index)
    dtype = np.dtype(dtype)
    arr = b2storage.allocate(
        tuple(int(s) for s in shape), tuple(layout_map), dtype, max(1, alignment_bytes // dtype.itemsize),
        tuple(aligned_index or (0,) * len(shape)),
    )  # fmt: skip
    return arr._base, arr


def _asarray(array, *, device=None):
    if gt_storage_utils.cp is None and (
        device == "gpu" or isinstance(array, b2storage.DeviceArray) or type(array).__module__.startswith("torch")
    ):
        if isinstance(array, b2storage.DeviceArray) or type(array).__module__.startswith("torch"):
            return array
        if hasattr(array, "__cuda_array_interface__"):
            return array
        return np.asarray(array)  # host data on its way into a device storage (from_array)
    return _orig_asarray(array, device=device)


def _cpu_copy(array):
    if isinstance(array, b2storage.DeviceArray) or type(array).__module__.startswith("torch"):
        return b2storage.cpu_copy(array)
    return _orig_cpu_copy(array)


def install_storage_hooks() -> None:
    gt_storage_utils.allocate_gpu = _allocate_gpu
    gt_storage_utils.asarray = _asarray
    gt_storage_utils.cpu_copy = _cpu_copy


B200_LAYOUT = gt_layout.LayoutInfo(
    alignment=b2storage.ALIGNMENT_ELEMENTS,
    device="gpu",
    layout_map=gt_layout.layout_maker_factory(b2storage.BASE_LAYOUT),
    is_optimal_layout=b2storage.is_optimal_layout,
)


# ---- run path ----------------------------------------------------------------------------------
class _ArrayProxy:
    """What `StencilObject._validate_args` needs from an array: shape / dtype / ndim / strides."""

    __slots__ = ("view", "shape", "dtype", "ndim", "strides")

    def __init__(self, view: runtime.ArrayView):
        self.view = view
        self.shape = view.shape
        self.dtype = view.dtype
        self.ndim = view.ndim
        self.strides = tuple(s * view.dtype.itemsize for s in view.strides)


class B200StencilObject(StencilObject):
    """StencilObject whose array handling does not need cupy.

    Only `_call_run`'s array extraction differs from the reference (stencil_object.py:531-612);
    origin normalisation, domain inference, validation and the per-call cache are inherited."""

    def _call_run(self, field_args, parameter_args, domain, origin, *, validate_args=True, exec_info=None):
        if exec_info is not None:
            exec_info["call_run_start_time"] = time.perf_counter()
        from gt4py.cartesian.definitions import LITERAL_INT_PRECISION, get_integer_type
        from gt4py.cartesian.frontend import gtscript_frontend

        lip = self.options.get("literal_int_precision", LITERAL_INT_PRECISION)
        for name, value in parameter_args.items():
            if type(value) in gtscript_frontend._ENUM_REGISTER.values():
                parameter_args[name] = get_integer_type(lip)(value.value)

        if any(a is not None and hostpipe.is_host_array(a) for a in field_args.values()):
            # arguments in HOST memory (an extension: the reference's GPU backends refuse them): the stand-alone mirror of
            # this stencil (same IR, options and cubin) stages them through device storages — hostpipe.host_call
            self._host_mirror()._call_run(field_args, parameter_args, domain, origin, validate_args=validate_args, exec_info=exec_info)
            if exec_info is not None:
                exec_info["call_run_end_time"] = time.perf_counter()
            return
        array_infos: Dict[str, Optional[ArgsInfo]] = {}
        for name, arg in field_args.items():
            if arg is None:
                array_infos[name] = None
                continue
            view = runtime.as_view(arg)
            dims = gt_storage_utils.get_dims(arg)
            if dims is not None:
                sorted_dims = [d for d in "IJK" if d in dims]
                sorted_dims += [str(d) for d in sorted(int(d) for d in dims if str(d).isdigit())]
                view = view.transpose([dims.index(d) for d in sorted_dims])
                dims = tuple(sorted_dims)
            array_infos[name] = ArgsInfo(
                array=_ArrayProxy(view), original_object=arg, dimensions=dims, device="gpu",
                origin=gt_storage_utils.get_origin(arg),
            )  # fmt: skip
        cache_key = _compute_domain_origin_cache_key(array_infos, parameter_args, domain, origin)
        if cache_key not in self._domain_origin_cache:
            origin = self._normalize_origins(array_infos, self.field_info, origin)
            if domain is None:
                domain = self._get_max_domain(array_infos, self.domain_info, self.field_info, origin)
            if validate_args:
                self._validate_args(array_infos, parameter_args, domain, origin)
            type(self)._domain_origin_cache[cache_key] = (domain, origin)
        else:
            domain, origin = type(self)._domain_origin_cache[cache_key]
        views = {n: (i.array.view if i is not None else None) for n, i in array_infos.items()}
        self.run(_domain_=domain, _origin_=origin, exec_info=exec_info, **views, **parameter_args)
        if exec_info is not None:
            exec_info["call_run_end_time"] = time.perf_counter()


    def _host_mirror(self):
        mirror = type(self).__dict__.get("_b200_host_mirror")
        if mirror is None:
            from .stencil import B200Stencil

            g = type(self).run.__globals__  # the generated module: IR file + code-generation options (B200ModuleGenerator)
            opts = json.loads(g["_B200_OPTS"])
            mirror = B200Stencil.__new__(B200Stencil)
            B200Stencil._init_from_compiled(mirror, get_compiled(g["_B200_IR"], g["_B200_OPTS"]), opts)
            type(self)._b200_host_mirror = mirror
        return mirror


_COMPILED: Dict[str, runtime.CompiledStencil] = {}


def get_compiled(ir_path: str, options_json: str) -> runtime.CompiledStencil:
    key = ir_path + "|" + options_json
    cs = _COMPILED.get(key)
    if cs is None:
        st, opts = b2ir.load_file(ir_path), json.loads(options_json)
        path = pathlib.Path(ir_path)
        # warm start: the cubin + launch plan persisted next to the module in .gt_cache (generate())
        cs = runtime.CompiledStencil.load(st, opts, path.parent, path.stem)
        if cs is None:
            cs = runtime.CompiledStencil(st, opts)
            try:
                cs.save(path.parent, path.stem)
            except OSError:
                pass  # read-only cache directory: keep the in-memory build
        cs._persist = (path.parent, path.stem)  # lazily specialised variants live next to the module as well
        _COMPILED[key] = cs
    return cs


DeviceTimer = runtime.DeviceTimer


def run_compiled(cs: runtime.CompiledStencil, domain, origin, exec_info, fields, params, device_sync: bool) -> None:
    """Body of the generated `run()`: native call + optional stream synchronisation
    (reference: gtc_common.py:157-163, 288-296).  With `exec_info`, the call is bracketed by device events:
    `exec_info["run_device_time"]` = seconds the kernels took on the device (`device_sync=True`), or
    `exec_info["b200_device_timer"].elapsed()` for asynchronous calls (synchronises when asked)."""
    timer = None
    if exec_info is not None:
        stream = runtime.current_stream_handle()
        timer = DeviceTimer()
        timer.start(stream)
    n = cs.run(fields, params, tuple(int(d) for d in domain), origin)
    if timer is not None:
        timer.stop(stream)
    if device_sync:
        lib = runtime.load_library()
        runtime.check(lib.b200_stream_synchronize(runtime.current_stream_handle()))
    if exec_info is not None:
        exec_info["b200_kernel_launches"] = n
        if device_sync:
            exec_info["run_device_time"] = timer.elapsed()
        else:
            exec_info["b200_device_timer"] = timer


class B200ModuleGenerator(BaseModuleGenerator):
    def generate_imports(self) -> str:
        return "import pathlib\nfrom gt4py_b200 import backend as _b200"

    def generate_base_class_name(self) -> str:
        return "_b200.B200StencilObject"

    def generate_module_members(self) -> str:
        ir_name = self.builder.backend.ir_file_name  # type: ignore[attr-defined]
        opts = json.dumps(self.builder.backend.codegen_options())  # type: ignore[attr-defined]
        return (
            f"_B200_IR = str(pathlib.Path(__file__).parent / {ir_name!r})\n"
            f"_B200_OPTS = {opts!r}\n"
        )

    def generate_implementation(self) -> str:
        gtir = self.builder.gtir
        from gt4py.cartesian.gtc import gtir as gtir_mod

        fields = [p.name for p in gtir.params if isinstance(p, gtir_mod.FieldDecl)]
        params = [p.name for p in gtir.params if isinstance(p, gtir_mod.ScalarDecl)]
        sync = self.builder.options.backend_opts.get("device_sync", True)
        fdict = ", ".join(f"{n}={n}" for n in fields)
        pdict = ", ".join(f"{n}={n}" for n in params)
        return (
            "_b200.run_compiled(_b200.get_compiled(_B200_IR, _B200_OPTS), _domain_, _origin_, exec_info, "
            f"dict({fdict}), dict({pdict}), {bool(sync)})"
        )


@gt_backend.register
class B200Backend(BaseBackend):
    """`backend="b200"`: hand-written-style sm_100a CUDA kernels behind gt4py's backend API."""

    name = "b200"
    options: ClassVar[dict[str, Any]] = {
        "device_sync": {"versioning": True, "type": bool},
        "oir_pipeline": {"versioning": True, "type": passes.OirPipeline},
        "strategy": {"versioning": True, "type": str},  # "auto" | "point"
        "fmad": {"versioning": True, "type": bool},
        "opt_level": {"versioning": True, "type": str},  # "0".."3" like gt:gpu (ints accepted too)
        "extra_opt_flags": {"versioning": True, "type": str},  # extra nvcc flags, space separated
        "debug_mode": {"versioning": True, "type": bool},
        "add_profile_info": {"versioning": True, "type": bool},  # accepted for gt:gpu compatibility (-lineinfo is always on)
        "clean": {"versioning": False, "type": bool},  # accepted for gt:gpu compatibility (nothing to clean: no build tree)
        "tile_j": {"versioning": True, "type": int},
        "warps": {"versioning": True, "type": int},
        "vector_width": {"versioning": True, "type": int},
        "prefetch": {"versioning": True, "type": int},
        "l2_prefetch": {"versioning": True, "type": int},
        "seq_cache": {"versioning": True, "type": bool},
        "seq_prefetch": {"versioning": True, "type": bool},
        "interior_loop": {"versioning": True, "type": object},  # True | "steady"
        "static_pitch": {"versioning": True, "type": int},
        "stcs": {"versioning": True, "type": bool},  # streaming (evict-first) stores
        "ldcs": {"versioning": True, "type": bool},  # streaming loads instead of the read-only path
        "min_blocks": {"versioning": True, "type": int},
        "tma": {"versioning": True, "type": int},  # bulk-async variant: shared-memory ring slots (trips) per warp
        "tma_rows": {"versioning": True, "type": int},
        "tma_mode": {"versioning": True, "type": str},  # "tensor" (cp.async.bulk.tensor) | "bulk" (one cp.async.bulk per row)
        "tma_smem_kb": {"versioning": True, "type": int},
        "uniform_task": {"versioning": True, "type": bool},  # task index from a lane-0 shuffle: uniform-datapath addressing
        "div_inv": {"versioning": True, "type": bool},  # divisions by launch-invariant divisors through the hoisted reciprocal (default on)
        "div_slow": {"versioning": True, "type": str},  # "call" | "inline": where the IEEE fallback of those divisions lives
        "period": {"versioning": True, "type": int},  # streaming kernels: force the unroll factor of the march loop
        "k_order": {"versioning": True, "type": object},  # "auto" | bool: level-fastest task order (kernels reading at K offsets)
        "col_smem": {"versioning": True, "type": object},  # bool: temporaries of fused sweeps in shared memory
        "col_smem_block": {"versioning": True, "type": object},  # (bx, by) threads of such a kernel
        "col_smem_kb": {"versioning": True, "type": int},  # shared-memory budget of a CTA for them
        "col_hints": {"versioning": True, "type": bool},  # column kernels: streaming loads / stores for read-once / write-only fields
        "seq_rotate": {"versioning": True, "type": object},  # "auto" | bool: look-ahead as a register ring instead of a shifting pipeline
        "seq_smem_pad": {"versioning": True, "type": int},  # unused dynamic shared memory per CTA (occupancy cap)
        "thin_tile_j": {"versioning": True, "type": int},  # J tile of one- and two-level sections
        "halo_lean": {"versioning": True, "type": bool},  # halo_wait kernels: select-based task decode + flag wait without time limit
        "halo_wait": {"versioning": True, "type": bool},  # multi-GPU: boundary tiles wait for the neighbours' pushed halo rows
        "row_pointers": {"versioning": True, "type": bool},
        "fuse_columns": {"versioning": True, "type": bool},
        "fuse_loops": {"versioning": True, "type": bool},
        "specialize": {"versioning": True, "type": str},  # "off" | "lazy": per-pitch kernels at first call
        "verbose": {"versioning": False, "type": bool},
    }
    storage_info: ClassVar[gt_layout.LayoutInfo] = B200_LAYOUT
    languages: ClassVar[dict] = {"computation": "cuda", "bindings": ["python"]}
    MODULE_GENERATOR_CLASS = B200ModuleGenerator
    compile_cubin: ClassVar[bool] = True

    @property
    def ir_file_name(self) -> str:
        caching = self.builder.caching
        return f"{caching.module_prefix}b200_ir{caching.module_postfix}.json"

    def codegen_options(self) -> Dict[str, Any]:
        keep = ("strategy", "fmad", "opt_level", "extra_opt_flags", "debug_mode", "tile_j", "warps", "verbose", "vector_width", "prefetch",
                "l2_prefetch", "seq_cache", "seq_prefetch", "interior_loop", "static_pitch", "specialize", "stcs", "ldcs", "min_blocks", "fuse_loops", "row_pointers", "fuse_columns", "tma", "tma_rows", "tma_mode", "tma_smem_kb", "halo_wait", "uniform_task", "div_inv", "div_slow", "period", "k_order", "col_smem", "col_smem_block", "col_smem_kb", "col_hints", "seq_rotate", "seq_smem_pad", "thin_tile_j", "halo_lean")  # fmt: skip
        return {k: v for k, v in self.builder.options.backend_opts.items() if k in keep}

    def lower(self) -> Dict[str, Any]:
        key = "b200:ir"
        if key not in self.builder.backend_data:
            base_oir = GTIRToOIR().visit(self.builder.gtir)
            pipeline = self.builder.options.backend_opts.get("oir_pipeline", from_oir.default_pipeline("staged"))
            st = from_oir.lower_oir(pipeline.run(base_oir))
            st.update(from_oir.args_data_to_ir(make_args_data_from_gtir(self.builder.gtir_pipeline)))
            st["name"] = self.builder.options.name
            self.builder.with_backend_data({key: st})
        return self.builder.backend_data[key]

    def generate(self):
        self.check_options(self.builder.options)
        build_info = self.builder.options.build_info
        t0 = time.perf_counter()
        st = self.lower()
        if build_info is not None:
            build_info["codegen_time"] = time.perf_counter() - t0
        if not self.builder.options._impl_opts.get("disable-code-generation", False):
            src_dir = self.builder.module_path.parent
            src_dir.mkdir(parents=True, exist_ok=True)
            b2ir.save_file(st, src_dir / self.ir_file_name)
            t1 = time.perf_counter()
            # generate + nvcc now (build errors surface at decoration time, like the reference)
            if self.compile_cubin:
                cs = runtime.CompiledStencil(st, self.codegen_options())
                cs.save(src_dir, pathlib.Path(self.ir_file_name).stem)  # <stem>.cubin / .plan.json / .cu
            if build_info is not None:
                build_info["build_time"] = time.perf_counter() - t1
        return self.make_module()

    @property
    def extra_cache_info(self) -> Dict[str, Any]:
        def md5(path):
            return hashlib.md5(path.read_bytes()).hexdigest() if path.exists() else ""

        path = self.builder.module_path.parent / self.ir_file_name
        return {**super().extra_cache_info, "b200_ir_md5": md5(path), "b200_cubin_md5": md5(path.with_suffix(".cubin"))}

    @property
    def extra_cache_validation_keys(self):
        """gt4py re-validates these against the stored cache info before reusing a cached module
        (reference: caching.py:214-265): a tampered / truncated IR or cubin forces a rebuild."""
        keys = super().extra_cache_validation_keys
        info = self.extra_cache_info
        keys += [k for k in ("b200_ir_md5", "b200_cubin_md5") if info[k]]
        return keys


install_storage_hooks()
