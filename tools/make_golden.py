"""Generate tests/golden/: lowered IR of every fixture stencil + outputs of the REFERENCE numpy backend.

Run in the build container only (needs /root/reference and tools/shims):

    PYTHONPATH=tools/shims:/root/reference/src:. GT_CACHE_ROOT=/tmp/gtcache python tools/make_golden.py

For every case in tools/stencil_defs.py:
  * tests/golden/ir/<case>.<variant>.json   — the b200 IR (gt4py frontend + OIR passes + from_oir)
  * tests/golden/<case>.npz                 — the written API fields after ONE call of the stencil
    compiled with the reference's own `backend="numpy"` on the seeded inputs of
    gt4py_b200.testing.make_case_data (seed 0); also seed 1 for the benchmark stencils.
It also asserts that the in-repo oracle (oracle/numpy_oracle.py) reproduces the reference bit for bit.
"""

from __future__ import annotations

import pathlib
import sys
import warnings

import numpy as np

warnings.filterwarnings("ignore")

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))

from gt4py.cartesian import gtscript  # noqa: E402

import stencil_defs  # noqa: E402
from gt4py_b200 import from_oir, ir as b2ir, testing  # noqa: E402
from oracle import numpy_oracle  # noqa: E402


def main(only=None):
    testing.IR_DIR.mkdir(parents=True, exist_ok=True)
    for name, case in stencil_defs.REGISTRY.items():
        if only and name not in only:
            continue
        irs = {}
        for variant in case["variants"]:
            st = from_oir.lower_definition(
                case["definition"], name=name, externals=case["externals"], variant=variant, **case["build"]
            )
            b2ir.save_file(st, testing.IR_DIR / f"{name}.{variant}.json")
            irs[variant] = st
        st = irs["default"]
        ref = gtscript.stencil(
            backend="numpy", definition=case["definition"], externals=case["externals"] or {}, name=name + "_ref", **case["build"]
        )
        out = {}
        for seed in (0, 1):
            fields, params, origins, domain = testing.make_case_data(st, name, seed=seed)
            ref_fields = {k: (v.copy() if v is not None else None) for k, v in fields.items()}
            ref(**ref_fields, **params, origin=origins, domain=domain)
            for variant, stv in irs.items():
                ofields = {k: (v.copy() if v is not None else None) for k, v in fields.items()}
                numpy_oracle.run(stv, ofields, params, domain, origins)
                for fname in testing.written_fields(st):
                    a, b = ref_fields[fname], ofields[fname]
                    if not np.array_equal(a, b, equal_nan=True):
                        bad = np.argwhere(~((a == b) | (np.isnan(a) & np.isnan(b)))) if a.dtype.kind == "f" else np.argwhere(a != b)
                        raise AssertionError(f"oracle != reference for {name}.{variant}:{fname} seed {seed}: {len(bad)} mismatches, first {bad[:3]}")
            for fname in testing.written_fields(st):
                out[f"seed{seed}.{fname}"] = ref_fields[fname]
        np.savez_compressed(testing.GOLDEN_DIR / f"{name}.npz", **out)
        print(f"{name}: ok ({', '.join(testing.written_fields(st))})")


if __name__ == "__main__":
    main(set(sys.argv[1:]) or None)
