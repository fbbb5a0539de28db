"""Every stencil of the REFERENCE's own definition library
(tests/cartesian_tests/integration_tests/multi_feature_tests/stencil_definitions.py — what its
test_code_generation.py::test_generation builds for every backend) is lowered through the b200
plug-in path, its generated CUDA is executed on the CPU emulator, and the result must equal the
reference numpy backend run on the same inputs.  Needs the reference tree (build container only)."""

import sys
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.needs_gt4py
REF_TESTS = "/root/reference/tests"


def _library():
    try:
        warnings.filterwarnings("ignore")
        if REF_TESTS not in sys.path:
            sys.path.insert(0, REF_TESTS)
        from cartesian_tests.integration_tests.multi_feature_tests import stencil_definitions as sd

        return sd
    except Exception:
        return None


_sd = _library()
NAMES = sorted(_sd.REGISTRY.names) if _sd is not None else []


@pytest.mark.parametrize("name", NAMES)
def test_reference_library_stencil(name):
    from gt4py.cartesian import gtscript

    from emu.emu import EmuStencil
    from gt4py_b200 import from_oir, testing
    from oracle import numpy_oracle

    defn, ext = _sd.REGISTRY[name], _sd.EXTERNALS_REGISTRY[name]
    st = from_oir.lower_definition(defn, name=name, externals=ext or None, variant="staged")
    fields, params, origins, domain = testing.make_case_data(st, name, domain=(9, 7, max(17, st["domain_info"]["min_k"])), seed=0)
    ref = gtscript.stencil(backend="numpy", definition=defn, externals=ext or {}, name=name + "_numpy_ref")
    rf = {k: (v.copy() if v is not None else None) for k, v in fields.items()}
    ref(**rf, **params, origin=origins, domain=domain)
    of = {k: (v.copy() if v is not None else None) for k, v in fields.items()}
    numpy_oracle.run(st, of, params, domain, origins)
    ef = {k: (v.copy() if v is not None else None) for k, v in fields.items()}
    EmuStencil(st, {}, name=f"reflib_{name}").run(ef, params, domain, origins)
    for f in testing.written_fields(st):
        np.testing.assert_array_equal(of[f], rf[f], err_msg=f"oracle vs reference numpy: {name}:{f}")
        np.testing.assert_allclose(ef[f], rf[f], rtol=1e-12, atol=0, err_msg=f"generated code vs reference numpy: {name}:{f}")
