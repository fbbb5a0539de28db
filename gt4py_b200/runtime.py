"""Python side of the C-ABI: ctypes binding of libgt4py_b200.so and the compiled-stencil object.

This is the `run()` half of the reference's generated module (reference:
backend/templates/stencil_module.py.in:160-169 -> backend/gtc_common.py:144-168
`pyext_module.run_computation(domain, field, origin, …, scalars…, exec_info)`): it turns the
already-normalised (arrays, origins, domain, scalars) of one stencil call into the plain-C structs
of include/gt4py_b200.h and enqueues the kernels.  There is no CPU path: without the shared library
or without a CUDA device every entry point raises.
"""

from __future__ import annotations

import ctypes
import os
import pathlib
import struct
import threading
from typing import Any, Dict, Optional, Sequence, Tuple

import numpy as np

from . import codegen, ir as b2ir, jit

_PKG = pathlib.Path(__file__).resolve().parent
LIB_PATH = _PKG / "libgt4py_b200.so"

EXPORTED_SYMBOLS = (
    "b200_abi_version", "b200_sizeof_field", "b200_last_error", "b200_device_info", "b200_stencil_load", "b200_stencil_unload",
    "b200_stencil_num_fields", "b200_stencil_scalars_size", "b200_stencil_num_kernels", "b200_stencil_kernel_name",
    "b200_stencil_run", "b200_stencil_run_halo", "b200_halo_push", "b200_halo_wait", "b200_stream_create", "b200_stream_create_priority", "b200_stream_destroy", "b200_stream_synchronize",
    "b200_event_create", "b200_event_destroy", "b200_event_record", "b200_stream_wait_event",
    "b200_event_elapsed_ms", "b200_comm_unique_id", "b200_comm_init", "b200_comm_destroy",
    "b200_halo_exchange", "b200_pack_2d", "b200_copy_box", "b200_relayout",
    "b200_graph_begin", "b200_graph_end", "b200_graph_num_nodes", "b200_graph_launch", "b200_graph_destroy",
)  # fmt: skip


class B200Push(ctypes.Structure):
    """b200_push_t (include/gt4py_b200.h): one box of rows pushed into a neighbour's memory"""

    _fields_ = [
        ("src", ctypes.c_void_p), ("dst", ctypes.c_void_p),
        ("row_bytes", ctypes.c_size_t), ("rows", ctypes.c_size_t), ("levels", ctypes.c_size_t),
        ("src_row_pitch", ctypes.c_size_t), ("src_level_pitch", ctypes.c_size_t),
        ("dst_row_pitch", ctypes.c_size_t), ("dst_level_pitch", ctypes.c_size_t),
    ]  # fmt: skip


class B200Field(ctypes.Structure):
    _fields_ = [
        ("data", ctypes.c_void_p),
        ("strides", ctypes.c_int64 * 7),
        ("origin", ctypes.c_int32 * 3),
        ("shape", ctypes.c_int32 * 3),
    ]


class B200Halo(ctypes.Structure):
    _fields_ = [
        ("send_lo", ctypes.c_void_p),
        ("recv_lo", ctypes.c_void_p),
        ("send_hi", ctypes.c_void_p),
        ("recv_hi", ctypes.c_void_p),
        ("bytes", ctypes.c_size_t),
    ]


class B200Error(RuntimeError):
    pass


_lib = None
_lib_lock = threading.Lock()


def load_library(build_if_missing: bool = True):
    """dlopen libgt4py_b200.so (building it in-tree first if it is missing and nvcc exists)."""
    global _lib
    if _lib is not None:  # fast path of every launch: no lock once loaded
        return _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not LIB_PATH.exists():
            if not build_if_missing:
                raise B200Error(f"b200: {LIB_PATH} is missing (run __graft_entry__.build())")
            jit.build_launcher()
        lib = ctypes.CDLL(str(LIB_PATH))
        vp, ci, cz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
        pvp = ctypes.POINTER(ctypes.c_void_p)
        pi = ctypes.POINTER(ctypes.c_int)
        sig = {
            "b200_abi_version": (ci, []),
            "b200_sizeof_field": (cz, []),
            "b200_last_error": (ctypes.c_char_p, []),
            "b200_device_info": (ci, [ci, pi, pi, pi, pi]),
            "b200_stencil_load": (ci, [vp, cz, ctypes.c_char_p, pvp]),
            "b200_stencil_unload": (ci, [vp]),
            "b200_stencil_num_fields": (ci, [vp]),
            "b200_stencil_scalars_size": (cz, [vp]),
            "b200_stencil_num_kernels": (ci, [vp]),
            "b200_stencil_kernel_name": (ctypes.c_char_p, [vp, ci]),
            "b200_stencil_run": (ci, [vp, ctypes.POINTER(B200Field), ci, vp, cz, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32), vp]),
            "b200_stencil_run_halo": (ci, [vp, ctypes.POINTER(B200Field), ci, vp, cz, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32),
                                          vp, vp, ctypes.c_uint64, vp]),
            "b200_halo_push": (ci, [ctypes.POINTER(B200Push), ci, ctypes.POINTER(ctypes.c_void_p), ci, ctypes.c_uint64, vp]),
            "b200_halo_wait": (ci, [vp, vp, ctypes.c_uint64, vp]),
            "b200_stream_create": (ci, [pvp]),
            "b200_stream_create_priority": (ci, [pvp, ci]),
            "b200_stream_destroy": (ci, [vp]),
            "b200_stream_synchronize": (ci, [vp]),
            "b200_event_create": (ci, [pvp]),
            "b200_event_destroy": (ci, [vp]),
            "b200_event_record": (ci, [vp, vp]),
            "b200_stream_wait_event": (ci, [vp, vp]),
            "b200_event_elapsed_ms": (ci, [vp, vp, ctypes.POINTER(ctypes.c_float)]),
            "b200_comm_unique_id": (ci, [vp]),
            "b200_comm_init": (ci, [pvp, vp, ci, ci]),
            "b200_comm_destroy": (ci, [vp]),
            "b200_halo_exchange": (ci, [vp, ctypes.POINTER(B200Halo), ci, ci, ci, vp]),
            "b200_pack_2d": (ci, [vp, cz, vp, cz, cz, cz, vp]),
            "b200_copy_box": (ci, [vp, cz, cz, vp, cz, cz, cz, cz, cz, vp]),
            "b200_relayout": (ci, [vp, vp, ci, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64), vp]),
            "b200_graph_begin": (ci, [vp]),
            "b200_graph_end": (ci, [vp, pvp]),
            "b200_graph_num_nodes": (ci, [vp]),
            "b200_graph_launch": (ci, [vp, vp]),
            "b200_graph_destroy": (ci, [vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.b200_abi_version() != 1 or lib.b200_sizeof_field() != ctypes.sizeof(B200Field):
            raise B200Error("b200: ABI mismatch between gt4py_b200 and libgt4py_b200.so (rebuild the launcher)")
        _lib = lib
        return lib


def check(rc: int) -> int:
    if rc < 0:
        msg = load_library().b200_last_error().decode(errors="replace")
        raise B200Error(f"b200 launcher error {rc}: {msg}")
    return rc


def device_info(device: int = 0) -> Dict[str, int]:
    lib = load_library()
    n, ma, mi, sms = (ctypes.c_int(0) for _ in range(4))
    check(lib.b200_device_info(device, ctypes.byref(n), ctypes.byref(ma), ctypes.byref(mi), ctypes.byref(sms)))
    return {"n_devices": n.value, "sm_major": ma.value, "sm_minor": mi.value, "n_sms": sms.value}


# ---- array views ------------------------------------------------------------------------------------
class ArrayView:
    """(pointer, shape, element strides, dtype) of a device array, whatever object it came from."""

    __slots__ = ("ptr", "shape", "strides", "dtype", "obj")

    def __init__(self, ptr, shape, strides, dtype, obj=None):
        self.ptr = int(ptr)
        self.shape = tuple(int(s) for s in shape)
        self.strides = tuple(int(s) for s in strides)
        self.dtype = np.dtype(dtype)
        self.obj = obj

    @property
    def ndim(self):
        return len(self.shape)

    def transpose(self, axes):
        return ArrayView(self.ptr, [self.shape[a] for a in axes], [self.strides[a] for a in axes], self.dtype, self.obj)


_TORCH_TO_NP = None


def as_view(obj) -> ArrayView:
    """Device array-like -> ArrayView.  Accepts DeviceArray, torch CUDA tensors and anything that
    exports __cuda_array_interface__ (cupy, numba, …).  Host arrays are rejected: the reference
    also refuses CPU arrays for GPU backends (storage/cartesian/utils.py:176-215)."""
    global _TORCH_TO_NP
    from .storage import DeviceArray

    if isinstance(obj, ArrayView):
        return obj
    if isinstance(obj, DeviceArray):
        return ArrayView(obj.data_ptr, obj.shape, obj.element_strides, obj.dtype, obj)
    mod = type(obj).__module__
    if mod.startswith("torch"):
        import torch

        if _TORCH_TO_NP is None:
            _TORCH_TO_NP = {
                torch.bool: np.dtype("bool"), torch.int8: np.dtype("int8"), torch.int16: np.dtype("int16"),
                torch.int32: np.dtype("int32"), torch.int64: np.dtype("int64"), torch.float32: np.dtype("float32"),
                torch.float64: np.dtype("float64"),
            }  # fmt: skip
        if not obj.is_cuda:
            raise TypeError("b200: torch tensor arguments must live on a CUDA device")
        return ArrayView(obj.data_ptr(), obj.shape, obj.stride(), _TORCH_TO_NP[obj.dtype], obj)
    cai = getattr(obj, "__cuda_array_interface__", None)
    if cai is not None:
        dtype = np.dtype(cai["typestr"])
        shape = tuple(cai["shape"])
        strides = cai.get("strides")
        if strides is None:
            acc, es = 1, []
            for s in reversed(shape):
                es.append(acc)
                acc *= s
            estrides = tuple(reversed(es))
        else:
            if any(s % dtype.itemsize for s in strides):
                raise ValueError("b200: strides must be multiples of the item size")
            estrides = tuple(s // dtype.itemsize for s in strides)
        return ArrayView(cai["data"][0], shape, estrides, dtype, obj)
    raise TypeError(f"b200: cannot use {type(obj)} as a device field (need a CUDA array)")


_stream_override = threading.local()


class DeviceTimer:
    """Device-side duration of one call: two events recorded on the launching stream around the kernels
    (the counterpart of the reference's `run_cpp_start_time` / `run_cpp_end_time`, gtc_common.py:83-99, which
    bracket the native call with a host clock — meaningless for asynchronous launches)."""

    _pool = threading.local()

    def __init__(self):
        self._lib = load_library()
        ev = getattr(DeviceTimer._pool, "events", None)
        if ev is None:
            ev = (ctypes.c_void_p(), ctypes.c_void_p())
            check(self._lib.b200_event_create(ctypes.byref(ev[0])))
            check(self._lib.b200_event_create(ctypes.byref(ev[1])))
            DeviceTimer._pool.events = ev
        self._ev = ev

    def start(self, stream) -> None:
        check(self._lib.b200_event_record(self._ev[0], stream))

    def stop(self, stream) -> None:
        check(self._lib.b200_event_record(self._ev[1], stream))

    def elapsed(self) -> float:
        """Seconds between the two events (synchronises on the second one)."""
        ms = ctypes.c_float(0.0)
        check(self._lib.b200_event_elapsed_ms(self._ev[0], self._ev[1], ctypes.byref(ms)))
        return float(ms.value) * 1e-3


def current_stream_handle() -> int:
    """Stream the launcher enqueues on when a call passes none: torch's current stream, unless a
    StencilGraph capture redirected this thread's launches to its capture stream."""
    s = getattr(_stream_override, "handle", None)
    if s is not None:
        return s
    import torch

    return int(torch.cuda.current_stream().cuda_stream)


def set_stream_override(handle: Optional[int]) -> None:
    _stream_override.handle = handle


#: what `specialize="lazy"` adds to the options at the first call with a given row pitch (besides `static_pitch`):
#: the winners of the device sweeps (profiles/README.md, r02): interior warps without per-lane predicates
LAZY_VARIANT: Dict[str, Any] = {"interior_loop": True}
#: default of the `specialize` option ("lazy" | "off"); GT4PY_B200_SPECIALIZE overrides it (the test-suites pin "off"
#: where they count kernels / compilations)
DEFAULT_SPECIALIZE = "lazy"


# ---- compiled stencil -----------------------------------------------------------------------------
class CompiledStencil:
    """Generated CUDA code + launch plan of one stencil, loaded lazily into the launcher."""

    def __init__(self, stencil_ir: Dict[str, Any], options: Optional[Dict[str, Any]] = None, *, name: Optional[str] = None,
                 artifacts: Optional[Dict[str, Any]] = None):  # fmt: skip
        self.ir = stencil_ir
        self.options = dict(options or {})
        self.name = name or stencil_ir["name"]
        if artifacts is not None:  # warm start: persisted plan + cubin, no code generation, no nvcc
            self.source, self.plan, self.cubin = artifacts.get("source", ""), artifacts["plan"], artifacts["cubin"]
            self.plan_text = codegen.plan_to_text(self.plan)
        else:
            self.source, self.plan = codegen.generate(stencil_ir, self.options)
            self.plan_text = codegen.plan_to_text(self.plan)
            self.cubin = jit.compile_cubin(self.source, self.options, name=codegen._cname(self.name), verbose=bool(self.options.get("verbose")))
        self.from_artifacts = artifacts is not None
        # specialize="lazy": at the first call with a given row pitch, compile (once, disk-cached) the
        # variant of the streaming kernels with that pitch as a compile-time constant + the interior
        # steady loop (codegen_stream.py) and use it for every later call with the same pitch
        self._specialize = str(self.options.get("specialize", os.environ.get("GT4PY_B200_SPECIALIZE", DEFAULT_SPECIALIZE)))
        self._special: Dict[int, "CompiledStencil"] = {}
        self._persist: Optional[Tuple[Any, str]] = None  # (.gt_cache directory, stem): where specialisations are kept too
        self._handle = None
        self._api = [f for f in self.plan["fields"] if f["kind"] == "api"]
        self._scalars = self.plan["scalars"]
        self._scal_struct = self._make_scalar_packer()
        self.last_launches = 0
        self._dom_cache: Dict[Any, Any] = {}  # domain / sub-box tuples -> ctypes arrays (the launch path is hot)

    def _make_scalar_packer(self):
        fmt_of = {"bool": "?", "int8": "b", "int16": "h", "int32": "i", "int64": "q", "float32": "f", "float64": "d"}
        fmt, pos = "<", 0
        for s in self._scalars:
            if s["offset"] > pos:
                fmt += f"{s['offset'] - pos}x"
            fmt += fmt_of[s["dtype"]]
            pos = s["offset"] + b2ir.ITEMSIZE[s["dtype"]]
        if self.plan["scalars_size"] > pos:
            fmt += f"{self.plan['scalars_size'] - pos}x"
        return struct.Struct(fmt)

    # ---- on-disk artefacts (warm starts skip code generation and nvcc) ------------------
    def save(self, directory, stem: str) -> Dict[str, str]:
        """Write `<stem>.cubin`, `<stem>.plan.json` (+ `<stem>.cu` for inspection) into `directory`
        (the stencil's folder in gt4py's .gt_cache).  Returns {file name: md5} for the cache info."""
        import hashlib
        import json

        d = pathlib.Path(directory)
        d.mkdir(parents=True, exist_ok=True)
        meta = {
            "generator": jit.generator_fingerprint(),
            "options": {k: v for k, v in sorted(self.options.items()) if isinstance(v, (str, int, float, bool, type(None)))},
            "ir": b2ir.fingerprint(self.ir),
            "plan": self.plan,
        }
        files = {f"{stem}.cubin": self.cubin, f"{stem}.plan.json": json.dumps(meta, sort_keys=True, indent=1).encode()}
        if self.source:
            files[f"{stem}.cu"] = self.source.encode()
        out = {}
        for fname, data in files.items():
            tmp = d / (fname + f".tmp{os.getpid()}")
            tmp.write_bytes(data)
            os.replace(tmp, d / fname)  # atomic: concurrent builders of the same stencil race benignly
            out[fname] = hashlib.md5(data).hexdigest()
        return out

    @classmethod
    def load(cls, stencil_ir, options, directory, stem: str, *, name: Optional[str] = None) -> Optional["CompiledStencil"]:
        """Reuse persisted artefacts if they were produced by this generator version from this IR
        with these options; None otherwise (the caller regenerates)."""
        import json

        d = pathlib.Path(directory)
        try:
            meta = json.loads((d / f"{stem}.plan.json").read_text())
            cubin = (d / f"{stem}.cubin").read_bytes()
        except (OSError, ValueError):
            return None
        opts = {k: v for k, v in sorted(dict(options or {}).items()) if isinstance(v, (str, int, float, bool, type(None)))}
        if meta.get("generator") != jit.generator_fingerprint() or meta.get("options") != opts or not cubin:
            return None
        if meta.get("ir") != b2ir.fingerprint(stencil_ir):
            return None
        src = d / f"{stem}.cu"
        art = {"plan": meta["plan"], "cubin": cubin, "source": src.read_text() if src.exists() else ""}
        return cls(stencil_ir, options, name=name, artifacts=art)

    @property
    def handle(self):
        if self._handle is None:
            lib = load_library()
            h = ctypes.c_void_p()
            buf = ctypes.create_string_buffer(self.cubin, len(self.cubin))
            check(lib.b200_stencil_load(buf, len(self.cubin), self.plan_text.encode(), ctypes.byref(h)))
            self._handle = h
        return self._handle

    def kernel_names(self):
        return [k["name"] for k in self.plan["kernels"]]

    def pack_scalars(self, params: Dict[str, Any]) -> bytes:
        vals = []
        for s in self._scalars:
            v = params.get(s["name"])
            if v is None:
                v = 0
            vals.append(bool(v) if s["dtype"] == "bool" else (int(v) if s["dtype"].startswith("int") else float(v)))
        return self._scal_struct.pack(*vals)

    def make_field_descs(self, views: Dict[str, Optional[ArrayView]], origins: Dict[str, Sequence[int]]):
        arr = (B200Field * max(1, len(self._api)))()
        for n, f in enumerate(self._api):
            v = views.get(f["name"])
            d = arr[n]
            if v is None:
                d.data = None
                continue
            org = origins[f["name"]]
            d.data = v.ptr
            ax = 0
            for a in range(3):
                if f["dims"][a]:
                    d.strides[a] = v.strides[ax]
                    d.origin[a] = int(org[ax])
                    d.shape[a] = v.shape[ax]
                    ax += 1
                else:
                    d.strides[a] = 0
                    d.origin[a] = 0
                    d.shape[a] = 1
            for dd in range(len(f["data_dims"])):
                d.strides[3 + dd] = v.strides[ax + dd]
        return arr

    def run(self, fields: Dict[str, Any], params: Dict[str, Any], domain: Sequence[int], origins: Dict[str, Sequence[int]],
            *, stream: Optional[int] = None, subbox: Optional[Sequence[int]] = None) -> int:
        """Enqueue one stencil application; returns the number of kernel launches."""
        views = {n: (as_view(a) if a is not None else None) for n, a in fields.items()}
        return self.run_views(views, self.pack_scalars(params), domain, origins, stream=stream, subbox=subbox)

    def run_views(self, views, scalars: bytes, domain, origins, *, stream=None, subbox=None) -> int:
        descs = self.make_field_descs(views, origins)
        return self.run_descs(descs, scalars, domain, stream=stream, subbox=subbox)

    def specialized_for(self, descs) -> "CompiledStencil":
        """The compiled variant to run for these field descriptors (self unless specialize="lazy")."""
        if self._specialize != "lazy" or "static_pitch" in self.options:
            return self
        if not any(k["kind"] == "stream" for k in self.plan["kernels"]):
            return self
        pitches = set()
        for n, f in enumerate(self._api):
            d = descs[n]
            if d.data and all(f["dims"]) and d.strides[0] == 1:
                pitches.add(int(d.strides[1]))
        if len(pitches) != 1:
            return self
        pitch = pitches.pop()
        cs = self._special.get(pitch)
        if cs is None:
            opts = {**LAZY_VARIANT, **self.options, "static_pitch": pitch, "specialize": "off"}
            if self._persist is not None:  # warm start: the specialised cubin of an earlier process
                cs = CompiledStencil.load(self.ir, opts, self._persist[0], f"{self._persist[1]}.p{pitch}", name=self.name)
            if cs is None:
                try:
                    cs = CompiledStencil(self.ir, opts, name=self.name)
                except Exception:  # a variant that does not apply to this stencil: keep the generic kernels
                    cs = self
                if cs is not self and self._persist is not None:
                    try:
                        cs.save(self._persist[0], f"{self._persist[1]}.p{pitch}")
                    except OSError:
                        pass
            self._special[pitch] = cs
        return cs

    def run_descs(self, descs, scalars: bytes, domain, *, stream=None, subbox=None, halo_wait=None) -> int:
        """halo_wait = (flag_lo address or 0, flag_hi address or 0, epoch): see b200_stencil_run_halo"""
        target = self.specialized_for(descs)
        if target is not self:
            n = target.run_descs(descs, scalars, domain, stream=stream, subbox=subbox, halo_wait=halo_wait)
            self.last_launches = n
            return n
        lib = load_library()
        domain = tuple(domain)
        dom = self._dom_cache.get(domain)
        if dom is None:
            if len(self._dom_cache) > 64:
                self._dom_cache.clear()
            dom = self._dom_cache[domain] = (ctypes.c_int32 * 3)(*[int(d) for d in domain])
        sb = None
        if subbox is not None:
            bkey = ("box", *subbox)
            sb = self._dom_cache.get(bkey)
            if sb is None:
                sb = self._dom_cache[bkey] = (ctypes.c_int32 * 4)(*[int(x) for x in subbox])
        if stream is None:
            stream = current_stream_handle()
        if halo_wait is not None:
            flo, fhi, epoch = halo_wait
            n = check(lib.b200_stencil_run_halo(self.handle, descs, len(self._api), scalars, len(scalars), dom, sb,
                                                ctypes.c_void_p(flo or None), ctypes.c_void_p(fhi or None), int(epoch), ctypes.c_void_p(stream)))  # fmt: skip
        else:
            n = check(lib.b200_stencil_run(self.handle, descs, len(self._api), scalars, len(scalars), dom, sb, ctypes.c_void_p(stream)))
        self.last_launches = n
        return n

    def __del__(self):
        try:
            if self._handle is not None and _lib is not None:
                _lib.b200_stencil_unload(self._handle)
        except Exception:
            pass
