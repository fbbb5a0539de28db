"""Shared helpers for the GPU parity tests: CUDA path vs oracle on the same seeded inputs."""

import numpy as np

from gt4py_b200 import storage, testing
from gt4py_b200.stencil import B200Stencil
from oracle import numpy_oracle

# cases whose arithmetic goes through libm transcendental functions: device libm and NumPy differ
# by a few ulp, everything else (+,-,*,/,sqrt, comparisons, casts, integer work) must be bit-exact
TOLERANCE = {
    "math_f64": dict(rtol=1e-12, atol=1e-13),
    "math_f32": dict(rtol=0, atol=0),
    "rounding_f64": dict(rtol=1e-11, atol=1e-13),
}


def compare(name, fname, got, ref):
    tol = TOLERANCE.get(name)
    if tol is None or (tol["rtol"] == 0 and tol["atol"] == 0):
        np.testing.assert_array_equal(got, ref, err_msg=f"{name}:{fname}")
    else:
        np.testing.assert_allclose(got, ref, err_msg=f"{name}:{fname}", **tol)


def run_case(name, variant="default", options=None, domain=None, seed=0, *, to_device=None, check_golden=False):
    st = testing.load_ir(name, variant)
    if domain is not None:  # respect the stencil's minimum K size (domain_info.min_sequential_axis_size)
        domain = (domain[0], domain[1], max(domain[2], int(st["domain_info"]["min_k"])))
    fields, params, origins, domain = testing.make_case_data(st, name, domain=domain, seed=seed)
    ref_fields = {k: (v.copy() if v is not None else None) for k, v in fields.items()}
    numpy_oracle.run(st, ref_fields, params, domain, origins)
    if to_device is None:
        to_device = lambda n, a: storage.from_array(a, aligned_index=origins[n])  # noqa: E731
    dev = {k: (to_device(k, v) if v is not None else None) for k, v in fields.items()}
    stencil = B200Stencil(st, options, name=f"{name}.{variant}")
    stencil(**dev, **params, origin=origins, domain=domain)
    results = {}
    for fname in st["field_info"]:
        if dev.get(fname) is None:
            continue
        got = storage.cpu_copy(dev[fname])
        results[fname] = got
        # written fields must match the oracle; read-only fields must be untouched
        compare(name, fname, got, ref_fields[fname])
    if check_golden:
        golden = np.load(testing.GOLDEN_DIR / f"{name}.npz")
        for fname in testing.written_fields(st):
            compare(name, fname, results[fname], golden[f"seed{seed}.{fname}"])
    return stencil, results
