"""gt4py_b200 — B200-native stencil-execution backend for gt4py.cartesian (`backend="b200"`).

Layout of the package (only what the hot path needs):
  ir.py / from_oir.py      stencil IR (a plain-dict restatement of OIR) and the gt4py OIR -> IR lowering
  codegen.py               point / column CUDA generator (always applicable)
  codegen_stream.py        streaming J-march generator for PARALLEL blocks (the fast path)
  codegen_column.py        K-march column generator with register k-caches for FORWARD / BACKWARD blocks
  jit.py                   nvcc -> sm_100a cubin, on-disk cache
  csrc/, ../include/       C-ABI launcher libgt4py_b200.so (kernel launch, scratch, NCCL halo exchange)
  runtime.py               ctypes binding, compiled-stencil object
  storage.py               gt4py.storage-compatible pitched device allocator (no cupy)
  stencil.py               stand-alone StencilObject mirror (runs from serialised IR, no gt4py needed)
  backend.py               the gt4py plug-in proper (registered when gt4py is importable)
  distributed.py           J-slab decomposition + halo exchange
  hostpipe.py / graph.py   host-resident K-slab pipeline; stencil sequences as CUDA graphs
  fuse.py                  cross-stencil fusion / temporal blocking (a call sequence as one stencil)
  tune_worker.py           sacrificial child process of the autotuner (`B200Stencil.autotune_isolated`)
"""

__version__ = "0.1.0"

def register() -> bool:
    """Register `backend="b200"` with gt4py now (for processes that made gt4py importable only after importing this
    package).  True when the plug-in is registered."""
    global HAVE_GT4PY
    try:
        import gt4py.cartesian  # noqa: F401

        from . import backend as _backend  # noqa: F401

        HAVE_GT4PY = True
    except ImportError:
        HAVE_GT4PY = False
    return HAVE_GT4PY


try:  # register backend="b200" with gt4py when the frontend is available
    import gt4py.cartesian  # noqa: F401

    from . import backend as _backend  # noqa: F401

    HAVE_GT4PY = True
except ImportError:  # stand-alone mode: serialised IR + launcher only
    HAVE_GT4PY = False
