"""-m gpu, last file of the suite on purpose: error path of a CUDA-graph capture (no device has run it yet)."""

import numpy as np
import pytest

from test_zz_graph_gpu import _setup

pytestmark = pytest.mark.gpu


def test_capture_before_first_call_fails_loudly():
    """scratch for temporaries cannot be allocated while capturing: the launcher says so"""
    from gt4py_b200 import runtime
    from gt4py_b200.graph import StencilGraph

    st, fields, origins, domain, frozen, mk = _setup("point")
    a, co, o = mk(fields["in_field"]), mk(fields["coeff"]), mk(np.zeros_like(fields["in_field"]))
    graph = StencilGraph()
    with pytest.raises(runtime.B200Error, match="before capturing"):
        with graph:
            frozen(in_field=a, coeff=co, out_field=o)
    frozen(in_field=a, coeff=co, out_field=o)  # the stream is usable again
    import torch

    torch.cuda.synchronize()
