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
            self._steps.append((frozen, views, ranges))
        self.s_in, self.s_out = torch.cuda.Stream(), torch.cuda.Stream()
        self._ev_in = [torch.cuda.Event() for _ in self.chunks]
        self._ev_k = [torch.cuda.Event() for _ in self.chunks]
        self._torch = torch
        self.h2d_bytes = sum(self.fields[n]._base.numel() * self.fields[n].itemsize for n in self.inputs)
        self.d2h_bytes = sum(self.fields[n]._base.numel() * self.fields[n].itemsize for n in self.outputs)

    def __call__(self, **kwargs) -> int:
        """kwargs: field name -> PinnedMirror (inputs are read, outputs are overwritten), scalar
        parameters by name.  Asynchronous: completion is ordered on the current stream."""
        torch = self._torch
        cur = torch.cuda.current_stream()
        host = {n: kwargs[n].flat for n in set(self.inputs) | set(self.outputs)}
        params = {n: kwargs[n] for n in self.stencil._param_names if n in kwargs}
        self.s_in.wait_stream(cur)  # device staging buffers are free once earlier work has finished
        launches = 0
        for c, (frozen, views, ranges) in enumerate(self._steps):
            with torch.cuda.stream(self.s_in):
                for n in self.inputs:
                    lo, hi = ranges[n]
                    self.fields[n]._base[lo:hi].copy_(host[n][lo:hi], non_blocking=True)
                self._ev_in[c].record(self.s_in)
            cur.wait_event(self._ev_in[c])
            launches += frozen(**views, **params)
            self._ev_k[c].record(cur)
            self.s_out.wait_event(self._ev_k[c])
            with torch.cuda.stream(self.s_out):
                for n in self.outputs:
                    lo, hi = ranges[n]
                    host[n][lo:hi].copy_(self.fields[n]._base[lo:hi], non_blocking=True)
        cur.wait_stream(self.s_out)
        return launches
