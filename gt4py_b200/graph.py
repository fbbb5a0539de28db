"""Stencil sequences as CUDA graphs: capture once, replay with one driver call.

The caller side of the hot path is a *sequence* of stencil calls per model time step (reference
caller: examples/cartesian/demo_burgers.ipynb cell 12 — three Runge-Kutta stages, copies and boundary
conditions per step).  The reference pays its whole Python call path (stencil_object.py:296-643,
10-50 us per call) for every one of them; `freeze()` removes the validation but not the
per-launch host work.  `StencilGraph` records whatever the enclosed calls enqueue through the C-ABI
launcher (kernels, halo exchanges, forked streams) into a CUDA graph:

    step = StencilGraph()
    with step:                       # nothing executes here, launches are captured
        rk_stage(**a, **p); rk_stage(**b, **p); copy(**c)
    for _ in range(n_steps):
        step.launch()                # one cudaGraphLaunch per time step

Kernel arguments (device addresses, scalars, domain) are frozen at capture time: swap buffers by
capturing one graph per buffer rotation, not by re-binding.  Stencils must be built with
`device_sync=False`.  Scratch for temporaries is owned per (stencil, stream): a capture allocates what it
needs, and the launcher keeps every buffer a captured graph points into alive until the stencil is unloaded.
"""

from __future__ import annotations

import ctypes
from typing import Optional

from . import runtime


class StencilGraph:
    def __init__(self, stream: Optional[int] = None):
        self._lib = runtime.load_library()
        self._stream = stream
        self._handle: Optional[ctypes.c_void_p] = None
        self._own_stream: Optional[ctypes.c_void_p] = None
        self._cap_stream = 0
        self._capturing = False

    def _stream_handle(self) -> int:
        return int(self._stream) if self._stream is not None else runtime.current_stream_handle()

    @property
    def stream(self) -> int:
        """Handle of the stream being captured (valid inside the `with` block) — pass it to code that
        forks other streams from the sequence (e.g. HaloExchanger events)."""
        return self._cap_stream

    def __enter__(self) -> "StencilGraph":
        if self._handle is not None:
            raise runtime.B200Error("b200: this StencilGraph already holds a captured sequence")
        s = self._stream_handle()
        if s == 0:
            # the legacy default stream cannot be captured: capture on a private stream and route
            # this thread's launches (calls that pass no stream) to it for the duration of the block
            if self._own_stream is None:
                h = ctypes.c_void_p()
                runtime.check(self._lib.b200_stream_create(ctypes.byref(h)))
                self._own_stream = h
            s = int(self._own_stream.value)
        self._cap_stream = s
        runtime.check(self._lib.b200_graph_begin(ctypes.c_void_p(s)))
        runtime.set_stream_override(s)
        self._capturing = True
        return self

    def __exit__(self, exc_type, exc, tb) -> bool:
        self._capturing = False
        runtime.set_stream_override(None)
        h = ctypes.c_void_p()
        rc = self._lib.b200_graph_end(ctypes.c_void_p(self._cap_stream), ctypes.byref(h))
        if exc_type is not None:
            if rc >= 0:
                self._lib.b200_graph_destroy(h)
            return False  # the capture is closed either way; propagate the caller's exception
        runtime.check(rc)
        self._handle = h
        return False

    @property
    def num_nodes(self) -> int:
        if self._handle is None:
            raise runtime.B200Error("b200: nothing captured yet")
        return runtime.check(self._lib.b200_graph_num_nodes(self._handle))

    def launch(self, stream: Optional[int] = None) -> None:
        if self._handle is None:
            raise runtime.B200Error("b200: nothing captured yet")
        s = int(stream) if stream is not None else self._stream_handle()
        runtime.check(self._lib.b200_graph_launch(self._handle, ctypes.c_void_p(s)))

    def close(self) -> None:
        if self._handle is not None:
            self._lib.b200_graph_destroy(self._handle)
            self._handle = None
        if self._own_stream is not None:
            self._lib.b200_stream_destroy(self._own_stream)
            self._own_stream = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
