"""`B200_EMULATE_DEVICE=1 pytest -m gpu`: run the GPU test-suite's LOGIC on the CPU (test infrastructure only).

The `-m gpu` tests can only be run on the B200 box, where a bug in a test (not in the product) would stop
the suite under `-x`.  This module stands a fake device under them: storages live in host memory (same pitched
layout), every `b200_stencil_run` is executed by the CPU emulator of the generated kernels (tests/emu/emu.py)
on the very pointers / strides / origins the launcher would receive, streams and events are no-ops.  It says
nothing about the device; it checks data flow, shapes, argument plumbing and assertions of the tests
themselves.  Tests that need real device behaviour (CUDA graphs, full-size domains, child processes) skip.
"""

from __future__ import annotations

import contextlib
import ctypes

import numpy as np

SKIP = ("test_no_cpu_fallback",)


class _Event:
    clock = 0.0

    def __init__(self, enable_timing=False):
        self.at = 0.0

    def record(self, stream=None):
        _Event.clock += 0.25
        self.at = _Event.clock

    def synchronize(self):
        pass

    def elapsed_time(self, other):
        return max(other.at - self.at, 0.01)


class _Stream:
    cuda_stream = 0

    def synchronize(self):
        pass

    def wait_stream(self, other):
        pass

    def wait_event(self, event):
        pass


def _array_from_desc(desc, f):
    dtype = np.dtype("bool" if f["dtype"] == "bool" else f["dtype"])
    dims = [a for a in range(3) if f["dims"][a]]
    shape = [int(desc.shape[a]) for a in dims] + [int(d) for d in f["data_dims"]]
    strides = [int(desc.strides[a]) for a in dims] + [int(desc.strides[3 + d]) for d in range(len(f["data_dims"]))]
    origin = [int(desc.origin[a]) for a in dims] + [0] * len(f["data_dims"])
    span = 1 + sum((n - 1) * s for n, s in zip(shape, strides))
    buf = (ctypes.c_char * (span * dtype.itemsize)).from_address(int(desc.data))
    arr = np.ndarray(shape, dtype, buffer=buf, strides=[s * dtype.itemsize for s in strides])
    return arr, tuple(origin)


_STATE = {"installed": False}


def installed() -> bool:
    return _STATE["installed"]


def install(mp) -> None:
    """mp: a pytest.MonkeyPatch"""
    import torch

    mp.setitem(_STATE, "installed", True)  # undone together with the patches (the previous value is recorded first)

    from gt4py_b200 import runtime, storage

    from .emu import EmuStencil

    mp.setattr(torch.cuda, "is_available", lambda: True)
    mp.setattr(torch.cuda, "set_device", lambda d: None)
    mp.setattr(torch.cuda, "current_device", lambda: 0)
    mp.setattr(torch.cuda, "synchronize", lambda *a: None)
    mp.setattr(torch.cuda, "current_stream", lambda *a: _Stream())
    mp.setattr(torch.cuda, "Event", _Event)
    mp.setattr(torch.cuda, "Stream", _Stream)
    mp.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    mp.setattr(torch.Tensor, "pin_memory", lambda self: self)
    mp.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    mp.setattr(storage, "_device", lambda device=None: torch.device("cpu"))

    def cai(self):  # host tensors pose as device arrays for objects that export __cuda_array_interface__ themselves
        np_dtype = np.dtype(str(self.dtype).replace("torch.", ""))
        return {"shape": tuple(self.shape), "typestr": np_dtype.str, "data": (self.data_ptr(), False),
                "strides": tuple(s * np_dtype.itemsize for s in self.stride()), "version": 3}  # fmt: skip

    mp.setattr(torch.Tensor, "__cuda_array_interface__", property(cai))

    real_as_view = runtime.as_view

    def as_view(obj):
        if type(obj).__module__.startswith("torch") and not isinstance(obj, runtime.ArrayView):
            np_dtype = np.dtype(str(obj.dtype).replace("torch.", ""))
            return runtime.ArrayView(obj.data_ptr(), obj.shape, obj.stride(), np_dtype, obj)
        return real_as_view(obj)

    mp.setattr(runtime, "as_view", as_view)

    capture = {"on": False, "calls": [], "graphs": {}}

    def run_descs(self, descs, scalars, domain, *, stream=None, subbox=None, halo_wait=None):
        if capture["on"]:  # CUDA-graph capture: nothing executes, the launches are recorded with frozen arguments
            frozen_descs = type(descs).from_buffer_copy(descs)
            capture["calls"].append((self, frozen_descs, bytes(scalars), tuple(domain), None if subbox is None else tuple(subbox)))
            n = sum(1 for st in self.specialized_for(descs).plan["steps"] if st["t"] == "launch")
            self.last_launches = n
            return n
        target = self.specialized_for(descs)
        emu = getattr(target, "_emu", None)
        if emu is None:
            emu = target._emu = EmuStencil(target.ir, target.options, name=target.name)
        fields, origins = {}, {}
        for n, f in enumerate(target._api):
            d = descs[n]
            if not d.data:
                fields[f["name"]] = None
                continue
            fields[f["name"]], origins[f["name"]] = _array_from_desc(d, f)
        vals = target._scal_struct.unpack(scalars) if target._scalars else ()
        params = {s["name"]: v for s, v in zip(target._scalars, vals)}
        target._emu_domain = tuple(int(x) for x in domain)
        before = emu.launches
        emu.run(fields, params, tuple(int(x) for x in domain), origins, subbox=tuple(subbox) if subbox is not None else None,
                halo_wait=tuple(int(x or 0) for x in halo_wait) if halo_wait is not None else (0, 0, 0))
        self.last_launches = emu.launches - before
        return self.last_launches

    mp.setattr(runtime.CompiledStencil, "run_descs", run_descs)

    real_lib = runtime.load_library()

    class Lib:
        def b200_stream_create(self, ref):
            ctypes.cast(ref, ctypes.POINTER(ctypes.c_void_p))[0] = 0x5000
            return 0

        def b200_stream_create_priority(self, ref, high):
            ctypes.cast(ref, ctypes.POINTER(ctypes.c_void_p))[0] = 0x5100
            return 0

        def b200_event_create(self, ref):
            ctypes.cast(ref, ctypes.POINTER(ctypes.c_void_p))[0] = 0x6000
            return 0

        def b200_event_elapsed_ms(self, start, stop, ref):
            ctypes.cast(ref, ctypes.POINTER(ctypes.c_float))[0] = 0.01
            return 0

        def b200_copy_box(self, dst, dpitch, dlevel_rows, src, spitch, slevel_rows, row_bytes, rows, levels, stream):
            dst, src = int(getattr(dst, "value", dst)), int(getattr(src, "value", src))
            for lev in range(int(levels)):
                for r in range(int(rows)):
                    ctypes.memmove(dst + (lev * dlevel_rows + r) * dpitch, src + (lev * slevel_rows + r) * spitch, int(row_bytes))
            return 0

        def b200_relayout(self, dst, src, itemsize, shape, ds, ss, stream):
            dst, src = int(getattr(dst, "value", dst)), int(getattr(src, "value", src))
            dt = np.dtype(f"u{itemsize}")
            shp = [int(shape[d]) for d in range(3)]

            def view(ptr, st):
                span = 1 + sum((n - 1) * abs(int(st[d])) for d, n in enumerate(shp))
                buf = (ctypes.c_char * (span * itemsize)).from_address(ptr)
                return np.ndarray(shp, dt, buffer=buf, strides=[int(st[d]) * itemsize for d in range(3)])

            view(dst, ds)[...] = view(src, ss)
            return 0

        def b200_graph_begin(self, stream):
            capture["on"], capture["calls"] = True, []
            return 0

        def b200_graph_end(self, stream, ref):
            capture["on"] = False
            handle = 0x7000 + len(capture["graphs"])
            capture["graphs"][handle] = capture["calls"]
            ctypes.cast(ref, ctypes.POINTER(ctypes.c_void_p))[0] = handle
            return 0

        def b200_graph_num_nodes(self, h):
            return sum(sum(1 for st in cs.plan["steps"] if st["t"] == "launch") for cs, *_ in capture["graphs"][h.value])

        def b200_graph_launch(self, h, stream):
            for cs, descs, scalars, domain, subbox in capture["graphs"][h.value]:
                cs.run_descs(descs, scalars, domain, subbox=subbox)
            return 0

        def b200_graph_destroy(self, h):
            return 0

        def __getattr__(self, name):
            if name.startswith(("b200_stream_", "b200_event_")):
                return lambda *a: 0
            if name.startswith(("b200_comm_", "b200_halo_", "b200_pack_")):
                raise AssertionError(f"{name} needs a real device")
            return getattr(real_lib, name)

    mp.setattr(runtime, "load_library", lambda *a, **k: Lib())

    # the autotune child (`python -m gt4py_b200.tune_worker spec.json`) runs in-process on the fake device
    import io
    import subprocess

    real_run = subprocess.run

    def fake_subprocess_run(cmd, *a, **kw):
        if not (isinstance(cmd, (list, tuple)) and "gt4py_b200.tune_worker" in cmd):
            return real_run(cmd, *a, **kw)
        from gt4py_b200 import tune_worker

        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            rc = tune_worker.main([cmd[-1]])
        return subprocess.CompletedProcess(cmd, rc, stdout=buf.getvalue(), stderr="")

    mp.setattr(subprocess, "run", fake_subprocess_run)
