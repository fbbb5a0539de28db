"""-m gpu: stencil sequences captured into a CUDA graph through the C-ABI (gt4py_b200/graph.py)."""

import numpy as np
import pytest

from gt4py_b200 import testing

pytestmark = pytest.mark.gpu


def _setup(strategy):
    """one geometry for every buffer (shape and origin of in_field), so that an output can be the next input"""
    from gt4py_b200 import storage
    from gt4py_b200.stencil import B200Stencil

    st = testing.load_ir("hdiff_f32", "staged")
    fields, params, origins, domain = testing.make_case_data(st, "hdiff_f32", domain=(150, 70, 6), seed=21)
    shape, org = fields["in_field"].shape, origins["in_field"]
    fields["coeff"] = np.random.default_rng(22).random(shape, dtype=np.float32)
    fields["out_field"] = np.zeros(shape, np.float32)
    origins = {n: org for n in ("in_field", "out_field", "coeff")}
    stencil = B200Stencil(st, {"strategy": strategy, "device_sync": False})
    frozen = stencil.freeze(origin=origins, domain=domain)
    mk = lambda a: storage.from_array(a, aligned_index=org)  # noqa: E731
    return st, fields, origins, domain, frozen, mk


@pytest.mark.parametrize("strategy", ["auto", "point"])
def test_graph_replay_equals_eager_sequence(strategy):
    """two chained applications (out1 = hdiff(in), out2 = hdiff(out1)): one graph launch == two calls,
    bit for bit, and a replay is idempotent; the point strategy adds scratch temporaries + 3 kernels/call"""
    import torch

    from gt4py_b200.graph import StencilGraph

    st, fields, origins, domain, frozen, mk = _setup(strategy)
    a, co = mk(fields["in_field"]), mk(fields["coeff"])
    e1, e2 = mk(np.zeros_like(fields["in_field"])), mk(np.zeros_like(fields["in_field"]))
    n = frozen(in_field=a, coeff=co, out_field=e1) + frozen(in_field=e1, coeff=co, out_field=e2)  # eager (also warm-up)
    torch.cuda.synchronize()
    g1, g2 = mk(np.zeros_like(fields["in_field"])), mk(np.zeros_like(fields["in_field"]))
    graph = StencilGraph()
    with graph:
        frozen(in_field=a, coeff=co, out_field=g1)
        frozen(in_field=g1, coeff=co, out_field=g2)
    torch.cuda.synchronize()
    assert graph.num_nodes == n
    assert float(g2.torch().abs().sum()) == 0.0  # capture does not execute
    for _ in range(2):
        graph.launch()
    torch.cuda.synchronize()
    assert torch.equal(g1.torch(), e1.torch()) and torch.equal(g2.torch(), e2.torch())
    # and the eager result is the oracle's
    from oracle import numpy_oracle

    ref = {k: v.copy() for k, v in fields.items()}
    numpy_oracle.run(st, ref, {}, domain, origins)
    np.testing.assert_array_equal(g1.get(), ref["out_field"])
    graph.close()
