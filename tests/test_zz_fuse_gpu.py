"""-m gpu: cross-stencil fusion (gt4py_b200/fuse.py) on the device — the fused kernel against the oracle
of the fused IR, and against the two separate calls it replaces."""

import numpy as np
import pytest

from gt4py_b200 import fuse, testing

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("domain", [(150, 70, 3), (33, 5, 2)])
@pytest.mark.parametrize("options", [None, {"strategy": "point"}, {"interior_loop": True, "specialize": "lazy"}])
def test_two_hdiff_steps_fused(domain, options):
    from gt4py_b200 import storage
    from gt4py_b200.stencil import B200Stencil
    from oracle import numpy_oracle

    st = testing.load_ir("hdiff_f32", "staged")
    f2 = fuse.repeat(st, 2, carry=("in_field", "out_field"))
    fields, params, origins, domain = testing.make_case_data(f2, "hdiff_f32", domain=domain, seed=41)
    ref = {k: v.copy() for k, v in fields.items()}
    numpy_oracle.run(f2, ref, params, domain, origins)
    dev = {k: storage.from_array(v, aligned_index=origins[k]) for k, v in fields.items()}
    B200Stencil(f2, options)(**dev, origin=origins, domain=domain)
    np.testing.assert_array_equal(dev["out_field"].get(), ref["out_field"])
    # == two device calls: step 1 on the domain grown by 2, step 2 on the domain
    one = B200Stencil(st, None)
    mid = storage.zeros(fields["in_field"].shape, np.float32, aligned_index=origins["in_field"])
    out = storage.from_array(fields["out_field"], aligned_index=origins["out_field"])
    grown = tuple(o - 2 if a < 2 else o for a, o in enumerate(origins["in_field"]))
    grown_c = tuple(o - 2 if a < 2 else o for a, o in enumerate(origins["coeff"]))
    one(dev["in_field"], mid, dev["coeff"], origin={"in_field": grown, "out_field": grown, "coeff": grown_c},
        domain=(domain[0] + 4, domain[1] + 4, domain[2]))  # fmt: skip
    one(mid, out, dev["coeff"], origin={"in_field": origins["in_field"], "out_field": origins["out_field"], "coeff": origins["coeff"]}, domain=domain)
    np.testing.assert_array_equal(out.get(), ref["out_field"])


def test_hdiff_then_upwind_fused_with_scalars():
    from gt4py_b200 import storage
    from gt4py_b200.stencil import B200Stencil
    from oracle import numpy_oracle

    a, b = testing.load_ir("hdiff_f32", "staged"), testing.load_ir("upwind5_f32", "staged")
    fused = fuse.compose("hdiff_upwind", [(a, {"out_field": "phi"}), (b, {})], intermediates=["phi"])
    fields, params, origins, domain = testing.make_case_data(fused, "upwind5_f32", domain=(200, 90, 3), seed=43)
    ref = {k: v.copy() for k, v in fields.items()}
    numpy_oracle.run(fused, ref, params, domain, origins)
    dev = {k: storage.from_array(v, aligned_index=origins[k]) for k, v in fields.items()}
    stencil = fuse.fuse_stencils("hdiff_upwind", [(a, {"out_field": "phi"}), (b, {})], intermediates=["phi"])
    stencil(**dev, **params, origin=origins, domain=domain)
    assert stencil.compiled.last_launches == 1
    np.testing.assert_array_equal(dev["out"].get(), ref["out"])
