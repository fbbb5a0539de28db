"""Sacrificial autotuning process: `python -m gt4py_b200.tune_worker <spec.json>`.

`B200Stencil.autotune_isolated` runs the candidate sweep here, in a child process with its own CUDA
context, on synthetic arguments that have exactly the caller's geometry (shape, element strides,
dtype, origin and the 256-byte phase of the base address, which decide the kernels' vector path and
the static row pitch).  A code-generation variant that faults or hangs takes this process down, not
the caller: the parent only ever launches the variants that completed here bit-identically to the
default one.  Prints one JSON line: {"tuned": [[candidate, ms], …], "rejected": […]}.
"""

from __future__ import annotations

import json
import sys

import numpy as np


class _Strided:
    """A device buffer with prescribed element strides and base-address phase (test/tuning data)."""

    def __init__(self, shape, strides, dtype, phase, gen):
        import torch

        from . import storage

        dev = storage._device()  # raises without a CUDA device: no CPU path
        dtype = np.dtype(dtype)
        span = 1 + sum((n - 1) * s for n, s in zip(shape, strides)) if all(n > 0 for n in shape) else 1
        nbytes = span * dtype.itemsize + 512
        self._raw = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        off = (phase - self._raw.data_ptr()) % 256
        tdt = getattr(torch, dtype.name if dtype.name != "bool" else "bool")
        flat = self._raw[off : off + span * dtype.itemsize].view(tdt)
        if dtype.kind == "f":
            flat.copy_(torch.rand(span, device=dev, dtype=tdt, generator=gen))
        elif dtype.kind == "b":
            flat.copy_(torch.rand(span, device=dev, generator=gen) < 0.5)
        else:
            flat.copy_(torch.randint(0, 4, (span,), device=dev, generator=gen).to(tdt))
        self.array = storage.DeviceArray(flat, 0, shape, strides, dtype)


def main(argv=None) -> int:
    argv = sys.argv[1:] if argv is None else argv
    spec = json.loads(open(argv[0]).read())
    import torch

    from .stencil import B200Stencil

    torch.cuda.set_device(int(spec.get("device", 0)))
    from . import storage

    gen = torch.Generator(device=storage._device())
    gen.manual_seed(1234)
    fields = {}
    for name, f in spec["fields"].items():
        fields[name] = None if f is None else _Strided(f["shape"], f["strides"], f["dtype"], int(f["phase"]), gen).array
    stencil = B200Stencil(spec["ir"], spec["options"], name=spec["name"])
    origin = {k: tuple(v) for k, v in spec["origin"].items()}
    tuned = stencil.autotune(fields, spec["params"], domain=tuple(spec["domain"]), origin=origin,
                             candidates=spec.get("candidates"), iters=int(spec.get("iters", 10)))  # fmt: skip
    torch.cuda.synchronize()
    print(json.dumps({"tuned": [[c, ms] for c, ms in tuned], "rejected": stencil.tune_rejected}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
