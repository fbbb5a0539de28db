"""-m gpu, last file of the suite on purpose: a CUDA-graph capture that is the FIRST call of a stencil with scratch
temporaries (scratch is owned per (stencil, stream) and may be allocated while capturing; ADVICE r1: a buffer that a
captured graph points into is never freed before the stencil is unloaded)."""

import numpy as np
import pytest

from test_zz_graph_gpu import _setup

pytestmark = pytest.mark.gpu


def test_capture_as_first_call_and_scratch_growth_after_capture():
    import torch

    from gt4py_b200 import storage
    from gt4py_b200.graph import StencilGraph
    from oracle import numpy_oracle

    st, fields, origins, domain, frozen, mk = _setup("point")  # point strategy: 3 kernels + scratch temporaries per call
    a, co, o = mk(fields["in_field"]), mk(fields["coeff"]), mk(np.zeros_like(fields["in_field"]))
    graph = StencilGraph()
    with graph:  # never called before: the capture allocates its own scratch
        frozen(in_field=a, coeff=co, out_field=o)
    torch.cuda.synchronize()
    assert float(o.torch().abs().sum()) == 0.0
    ref = {k: v.copy() for k, v in fields.items()}
    numpy_oracle.run(st, ref, {}, domain, origins)
    graph.launch()
    torch.cuda.synchronize()
    np.testing.assert_array_equal(o.get(), ref["out_field"])

    # the same stencil on a much larger domain, on the capture stream's sibling AND on the current stream: scratch
    # grows; the captured graph must keep working on the buffer it was captured with (use-after-free before the fix)
    stencil = frozen.stencil_object
    big = (domain[0] * 2, domain[1] * 3, domain[2])
    bf, bp, borg, big = __import__("gt4py_b200").testing.make_case_data(st, "hdiff_f32", domain=big, seed=5)
    bdev = {k: storage.from_array(v, aligned_index=borg[k]) for k, v in bf.items()}
    stencil(**bdev, origin=borg, domain=big)
    torch.cuda.synchronize()
    bref = {k: v.copy() for k, v in bf.items()}
    numpy_oracle.run(st, bref, bp, big, borg)
    np.testing.assert_array_equal(bdev["out_field"].get(), bref["out_field"])
    o.fill(0)
    graph.launch()
    torch.cuda.synchronize()
    np.testing.assert_array_equal(o.get(), ref["out_field"])
    graph.close()
