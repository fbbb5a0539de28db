"""Host-resident call path: K-slab pipelining of H2D copy / stencil / D2H copy.

The reference's host-side entry is `gt4py.storage.from_array(host)` (a synchronous H2D copy,
storage/cartesian/interface.py:323-325) -> `stencil(...)` -> `np.asarray(out)` (D2H): three serial
phases, PCIe idle while the kernel runs and the two DMA directions never active together.  For a
PARALLEL stencil without vertical dependencies the K levels are independent, and with the backend's
(2,1,0) layout (K outermost) a slab of levels is ONE contiguous range of the pitched buffer.  So the
call is cut into `n_chunks` K slabs and software-pipelined over three streams:

    copy-in stream : H2D slab c+1 of every input      (PCIe, host -> device)
    compute stream : stencil on slab c                (HBM)
    copy-out stream: D2H slab c-1 of every output     (PCIe, device -> host)

Both DMA engines and the SMs are busy at the same time; the end-to-end time tends to
max(H2D bytes, D2H bytes) / PCIe bandwidth instead of their sum plus the kernel.

Host buffers are pinned mirrors with the *same* pitched layout as the device storage
(`PinnedMirror`), so every transfer is one contiguous DMA.  No CPU compute path: the stencil always
runs through the C-ABI launcher on the device.
"""

from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import storage as b2storage


def pipeline_supported(stencil_ir: Dict[str, Any]) -> Optional[str]:
    """None when K slabs of the domain are independent, else the reason they are not."""
    for loop in stencil_ir["loops"]:
        if loop["order"] != "parallel":
            return f"{loop['order']} computation (vertical dependency)"
        for sec in loop["sections"]:
            if [list(b) for b in sec["interval"]] != [["start", 0], ["end", 0]]:
                return "computation restricted to a vertical interval"
    for name, fi in stencil_ir["field_info"].items():
        if fi is None or fi["access"] == "NONE":
            continue
        if "K" not in fi["axes"]:
            return f"field {name} has no K axis"
        if tuple(fi["boundary"][2]) != (0, 0):
            return f"field {name} is read at a K offset"
    from . import ir as b2ir

    bad: List[str] = []

    def visit(e):
        if e["t"] == "field" and isinstance(e["off"], dict):
            bad.append(f"variable / absolute K access of {e['name']}")
        elif e["t"] == "iter" and e["axis"] == "K":
            bad.append("K index access")

    for *_ignored, he in b2ir.iter_hes(stencil_ir):
        b2ir.walk_exprs(he["body"], visit)
    return bad[0] if bad else None


def plan_chunks(nk: int, n_chunks: int) -> List[Tuple[int, int]]:
    """Split [0, nk) into at most n_chunks contiguous, nearly equal, non-empty level ranges."""
    n = max(1, min(int(n_chunks), int(nk)))
    base, rem = divmod(int(nk), n)
    out, lo = [], 0
    for c in range(n):
        hi = lo + base + (1 if c < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def slab_range(numel: int, offset: int, stride_k: int, k_origin: int, k0: int, k1: int, first: bool, last: bool) -> Tuple[int, int]:
    """Element range of the flat pitched buffer that holds levels [k0, k1) of the compute domain
    (plus everything before the first / after the last slab, so the union covers the buffer)."""
    lo = 0 if first else offset + (k_origin + k0) * stride_k
    hi = numel if last else offset + (k_origin + k1) * stride_k
    return max(0, lo), min(numel, hi)


class PinnedMirror:
    """Pinned host buffer with the pitched layout of a DeviceArray (`array` is a strided numpy view)."""

    def __init__(self, dev: "b2storage.DeviceArray", src: Optional[np.ndarray] = None):
        import torch

        self.dev = dev
        self.flat = torch.empty(dev._base.numel(), dtype=dev._base.dtype).pin_memory()
        self._view = torch.as_strided(self.flat, dev.shape, dev.element_strides, self.flat.storage_offset() + dev._offset)
        if src is not None:
            self._view.copy_(torch.from_numpy(np.ascontiguousarray(src)))

    @property
    def array(self) -> np.ndarray:
        return self._view.numpy()

    @property
    def nbytes(self) -> int:
        return int(self.flat.numel()) * self.flat.element_size()


class HostPipeline:
    """`pipe(in_field=mirror, …)`: one stencil application from pinned host buffers to pinned host
    buffers, K-slab pipelined.  `fields`: name -> DeviceArray (device staging, allocated with
    gt4py_b200.storage so that K is the outermost axis); reused by every call."""

    def __init__(self, stencil, fields: Dict[str, Any], *, origin: Dict[str, Sequence[int]], domain: Sequence[int],
                 n_chunks: int = 8):  # fmt: skip
        import torch

        why = pipeline_supported(stencil.ir)
        if why is not None:
            raise ValueError(f"b200 host pipeline: K slabs are not independent for {stencil.name}: {why}")
        self.stencil = stencil
        self.domain = tuple(int(d) for d in domain)
        self.fields = dict(fields)
        self.chunks = plan_chunks(self.domain[2], n_chunks)
        self.inputs, self.outputs = [], []
        for name, fi in stencil.field_info.items():
            if fi is None or fi.access == "NONE":
                continue
            dev = self.fields[name]
            if not isinstance(dev, b2storage.DeviceArray) or dev.ndim != 3:
                raise ValueError(f"b200 host pipeline: field {name} must be a 3-D gt4py_b200.storage array")
            si, sj, sk = dev.element_strides
            if not (si == 1 and sk >= sj * dev.shape[1]):
                raise ValueError(f"b200 host pipeline: field {name} is not in the (2,1,0) layout (K outermost)")
            if fi.access in ("READ", "READ_WRITE"):
                self.inputs.append(name)
            if fi.access in ("WRITE", "READ_WRITE"):
                self.outputs.append(name)
        # per chunk: frozen stencil on the K-sliced views + the flat ranges to move
        self._steps = []
        nch = len(self.chunks)
        for c, (k0, k1) in enumerate(self.chunks):
            views, org, ranges = {}, {}, {}
            for name in self.inputs + [n for n in self.outputs if n not in self.inputs]:
                dev = self.fields[name]
                o = tuple(int(x) for x in origin[name])
                views[name] = dev[:, :, o[2] + k0 : o[2] + k1]
                org[name] = (o[0], o[1], 0)
                ranges[name] = slab_range(dev._base.numel(), dev._offset, dev.element_strides[2], o[2], k0, k1, c == 0, c == nch - 1)
            frozen = stencil.freeze(origin=org, domain=(self.domain[0], self.domain[1], k1 - k0))
            self._steps.append((frozen, views, ranges, (k0, k1)))
        self._origin = {n: tuple(int(x) for x in origin[n]) for n in self.fields}
        self.s_in, self.s_out = torch.cuda.Stream(), torch.cuda.Stream()
        self._ev_in = [torch.cuda.Event() for _ in self.chunks]
        self._ev_k = [torch.cuda.Event() for _ in self.chunks]
        self._torch = torch
        self.h2d_bytes = sum(self.fields[n]._base.numel() * self.fields[n].itemsize for n in self.inputs)
        self.d2h_bytes = sum(int(np.prod(self.domain)) * self.fields[n].itemsize for n in self.outputs)

    def __call__(self, **kwargs) -> int:
        """kwargs: field name -> PinnedMirror (inputs are read, outputs are overwritten), scalar
        parameters by name.  Asynchronous: completion is ordered on the current stream."""
        torch = self._torch
        cur = torch.cuda.current_stream()
        host = {n: _flat_of(kwargs[n]) for n in set(self.inputs) | set(self.outputs)}
        params = {n: kwargs[n] for n in self.stencil._param_names if n in kwargs}
        self.s_in.wait_stream(cur)  # device staging buffers are free once earlier work has finished
        launches = 0
        for c, (frozen, views, ranges, (k0, k1)) in enumerate(self._steps):
            with torch.cuda.stream(self.s_in):
                for n in self.inputs:
                    lo, hi = ranges[n]
                    self.fields[n]._base[lo:hi].copy_(host[n][lo:hi], non_blocking=True)
                self._ev_in[c].record(self.s_in)
            cur.wait_event(self._ev_in[c])
            launches += frozen(**views, **params)
            self._ev_k[c].record(cur)
            self.s_out.wait_event(self._ev_k[c])
            for n in self.outputs:  # D2H of the compute domain only (strided DMA): the caller's halo is not touched
                dev = self.fields[n]
                copy_domain_box(_array_ptr(host[n], dev._offset), _array_ptr(dev._base, dev._offset), dev, self._origin[n],
                                self.domain, k0, k1, int(self.s_out.cuda_stream))  # fmt: skip
        cur.wait_stream(self.s_out)
        return launches


def copy_domain_box(dst_ptr: int, src_ptr: int, dev: "b2storage.DeviceArray", origin, domain, k0: int, k1: int, stream: int) -> None:
    """Copy levels [k0, k1) of the COMPUTE DOMAIN of a pitched 3-D storage between two buffers of identical layout
    (`dst_ptr` / `src_ptr` = addresses of array element [0,0,0] on either side) with one strided DMA (b200_copy_box):
    what a stencil wrote travels back to the host, the halo of the caller's array stays untouched."""
    import ctypes

    from . import runtime

    si, sj, sk = dev.element_strides
    item = dev.itemsize
    if si != 1 or sk % sj:
        raise ValueError("b200 host path: storage is not in the pitched (2,1,0) layout")
    off = (int(origin[0]) * si + int(origin[1]) * sj + (int(origin[2]) + k0) * sk) * item
    runtime.check(runtime.load_library().b200_copy_box(
        ctypes.c_void_p(dst_ptr + off), sj * item, sk // sj, ctypes.c_void_p(src_ptr + off), sj * item, sk // sj,
        int(domain[0]) * item, int(domain[1]), int(k1 - k0), ctypes.c_void_p(stream)))  # fmt: skip


def _array_ptr(flat, offset_elems: int) -> int:
    return int(flat.data_ptr()) + int(offset_elems) * flat.element_size()


def _flat_of(obj):
    flat = getattr(obj, "_b200_flat", None)  # storage.HostArray
    return flat if flat is not None else obj.flat  # PinnedMirror


# ---- StencilObject.__call__ with HOST arrays ---------------------------------------------------------------------------
def is_host_array(obj) -> bool:
    """NumPy arrays (incl. storage.HostArray) and other objects that only export the host array interface."""
    if isinstance(obj, np.ndarray):
        return True
    if isinstance(obj, b2storage.DeviceArray) or hasattr(obj, "__cuda_array_interface__"):
        return False
    if type(obj).__module__.startswith("torch"):
        return not obj.is_cuda
    return hasattr(obj, "__array_interface__")


def _host_view(arg):
    from . import runtime

    arr = arg.numpy() if type(arg).__module__.startswith("torch") else np.asarray(arg)
    if any(s % arr.itemsize for s in arr.strides):
        raise ValueError("b200: strides must be multiples of the item size")
    return arr, runtime.ArrayView(arr.ctypes.data, arr.shape, tuple(s // arr.itemsize for s in arr.strides), arr.dtype, arg)


def default_host_chunks(options, nk: int) -> int:
    """K slabs of a host-array call: option `host_chunks`, else GT4PY_B200_HOST_CHUNKS, else HOST_CHUNKS.  The call returns
    when the result is in host memory, so every call pays the pipeline's fill (H2D of the first slab) and drain (D2H of the
    last one): thinner slabs shorten both, while the per-slab launches stay hidden behind the PCIe transfers."""
    import os

    n = options.get("host_chunks") or os.environ.get("GT4PY_B200_HOST_CHUNKS") or HOST_CHUNKS
    return max(1, min(int(n), int(nk)))


HOST_CHUNKS = 10


def host_call(stencil, field_args, parameter_args, domain, origin, *, validate_args=True, exec_info=None) -> None:
    """One stencil application on arguments that live in HOST memory — an extension over the reference, whose GPU
    backends refuse CPU arrays (storage/cartesian/utils.py:176-215): code written for `backend="numpy"` with plain
    NumPy arrays runs unchanged on `backend="b200"`.

    The fields are staged through cached device storages in the backend's layout: H2D of every field the stencil
    reads, the kernels, D2H of every field it writes; the call returns when the results are in host memory.  When all
    fields are `storage.HostArray`s (pinned, same pitched layout) and the K levels are independent, the three phases
    are software-pipelined over K slabs (HostPipeline); any other host array takes one pageable copy + a device-side
    re-layout per field.  There is no CPU compute path: the stencil always runs on the device."""
    import torch

    from . import stencil as b2stencil

    host = {n: _host_view(a) for n, a in field_args.items() if a is not None and is_host_array(a)}
    infos = {}
    for n, a in field_args.items():
        if a is None:
            infos[n] = None
        elif n in host:
            infos[n] = b2stencil._ArgInfo(host[n][1], getattr(a, "__gt_origin__", None), None)
        else:
            infos[n] = b2stencil.extract_array_infos({n: a})[n]
    origin = stencil._normalize_origins(infos, origin)
    if domain is None:
        domain = stencil._get_max_domain(infos, origin)
    domain = tuple(int(d) for d in domain)
    if validate_args:
        stencil._validate_args(infos, parameter_args, domain, origin)
    cache = stencil.__dict__.setdefault("_host_staging", {})
    dev_args = dict(field_args)
    for n, (arr, _view) in host.items():
        fi = stencil.field_info[n]
        dims = [a for a in "IJK" if a in fi.axes] + [str(d) for d in range(len(fi.data_dims))]
        org = tuple(int(o) for o in origin[n]) + (0,) * (arr.ndim - len(origin[n]))
        key = (n, arr.shape, arr.dtype.str, org)
        dev = cache.get(key)
        if dev is None:
            if len(cache) > 4 * max(1, len(field_args)):
                cache.clear()
            dev = cache[key] = b2storage.empty(arr.shape, arr.dtype, aligned_index=org, dimensions=dims)
        dev_args[n] = dev
    reads = [n for n in host if stencil.field_info[n].access in ("READ", "READ_WRITE")]
    writes = [n for n in host if stencil.field_info[n].access in ("WRITE", "READ_WRITE")]

    def same_layout(n):
        a = field_args[n]
        return getattr(a, "_b200_flat", None) is not None and a._b200_layout == b2storage.layout_signature(dev_args[n])

    used = [n for n, fi in stencil.field_info.items() if fi is not None and fi.access != "NONE" and field_args.get(n) is not None]
    launches, path = None, "serial"
    if all(n in host and same_layout(n) for n in used) and pipeline_supported(stencil.ir) is None and domain[2] >= 2:
        pkey = ("pipeline", tuple((n, id(dev_args[n])) for n in used), domain, tuple(sorted((n, tuple(origin[n])) for n in used)))
        pipe = cache.get(pkey)
        if pipe is None:
            try:
                pipe = HostPipeline(stencil, {n: dev_args[n] for n in used}, origin=origin, domain=domain,
                                    n_chunks=default_host_chunks(stencil.backend_options, domain[2]))  # fmt: skip
            except ValueError:
                pipe = False  # e.g. a field that is not 3-D: serial path
            cache[pkey] = pipe
        if pipe:
            launches, path = pipe(**{n: field_args[n] for n in used}, **parameter_args), "pipeline"
    if launches is None:
        for n in reads:
            if same_layout(n):
                dev_args[n]._base.copy_(field_args[n]._b200_flat, non_blocking=True)
            else:
                dev_args[n][...] = host[n][0]
        views = {n: (b2stencil.extract_array_infos({n: a})[n].view if a is not None else None) for n, a in dev_args.items()}
        launches = stencil.compiled.run_views(views, stencil.compiled.pack_scalars(parameter_args), domain, origin)
        from . import runtime

        for n in writes:
            dev = dev_args[n]
            if same_layout(n) and dev.ndim == 3:
                copy_domain_box(_array_ptr(field_args[n]._b200_flat, dev._offset), _array_ptr(dev._base, dev._offset), dev, origin[n],
                                domain, 0, domain[2], runtime.current_stream_handle())  # fmt: skip
    torch.cuda.current_stream().synchronize()
    if path == "serial":
        for n in writes:
            dev = dev_args[n]
            if not (same_layout(n) and dev.ndim == 3):
                # the compute domain of the field (axes it has), every data-dimension entry
                fi = stencil.field_info[n]
                spatial = [d for d, m in zip(domain, fi.domain_mask) if m]
                box = tuple(slice(int(o), int(o) + int(d)) for o, d in zip(origin[n], spatial))
                host[n][0][box] = dev[box].get()
    if exec_info is not None:
        exec_info["b200_kernel_launches"] = launches
        exec_info["b200_host_path"] = path
