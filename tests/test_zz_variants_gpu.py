"""-m gpu: every code-generation variant the autotuner may pick is bit-identical to the default one
(and the default one to the oracle) on the device, in the storage layout, incl. odd extents."""

import numpy as np
import pytest

from gt4py_b200 import testing

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,domain", [("hdiff_f32", (389, 200, 3)), ("upwind5_f32", (262, 150, 2)), ("hdiff_f64", (130, 131, 2))])
def test_autotune_candidates_are_bit_identical(name, domain):
    import torch

    from gt4py_b200 import storage
    from gt4py_b200.stencil import B200Stencil
    from oracle import numpy_oracle

    st = testing.load_ir(name, "staged")
    fields, params, origins, domain = testing.make_case_data(st, name, domain=domain, seed=31)
    ref = {k: v.copy() for k, v in fields.items()}
    numpy_oracle.run(st, ref, params, domain, origins)
    dev = {k: storage.from_array(v, aligned_index=origins[k]) for k, v in fields.items()}
    stencil = B200Stencil(st, {"device_sync": False})
    cands = [c for c in B200Stencil.DEFAULT_CANDIDATES if not ({"tile_j", "warps", "l2_prefetch", "min_blocks"} & set(c))]
    tuned = stencil.autotune(dev, params, domain=domain, origin=origins, iters=2, candidates=cands)
    assert stencil.tune_rejected == [], stencil.tune_rejected
    assert len(tuned) >= 6
    assert any("static_pitch" in c for c, _ in tuned) and any(c.get("interior_loop") for c, _ in tuned)
    # the winner, through the public call
    for fname in testing.written_fields(st):
        dev[fname].fill(0)
    stencil(**dev, **params, origin=origins, domain=domain)
    torch.cuda.synchronize()
    for fname in testing.written_fields(st):  # (the tuner zeroed the whole output storage: compare the compute domain)
        box = tuple(slice(o, o + d) for o, d in zip(origins[fname], domain))
        np.testing.assert_array_equal(dev[fname].get()[box], ref[fname][box], err_msg=f"{name}:{fname} {stencil.backend_options}")


def test_isolated_autotune_adopts_a_verified_variant():
    """the sweep in a sacrificial child process (gt4py_b200/tune_worker.py) on synthetic arguments of the
    caller's geometry; the parent adopts the child's winner only after a bit-for-bit check on the real data"""
    import torch

    from gt4py_b200 import storage
    from gt4py_b200.stencil import B200Stencil
    from oracle import numpy_oracle

    name, domain = "hdiff_f32", (200, 130, 3)
    st = testing.load_ir(name, "staged")
    fields, params, origins, domain = testing.make_case_data(st, name, domain=domain, seed=33)
    ref = {k: v.copy() for k, v in fields.items()}
    numpy_oracle.run(st, ref, params, domain, origins)
    dev = {k: storage.from_array(v, aligned_index=origins[k]) for k, v in fields.items()}
    stencil = B200Stencil(st, {"device_sync": False})
    cands = [{}, {"interior_loop": True, "static_pitch": "auto"}, {"static_pitch": "auto"}, {"tile_j": 32}]
    tuned = stencil.autotune_isolated(dev, params, domain=domain, origin=origins, candidates=cands, iters=3, timeout=240)
    assert len(tuned) == 4 and stencil.tune_rejected == []
    assert {k: v for k, v in stencil.backend_options.items() if k != "device_sync"} in [c for c, _ in tuned]
    for fname in testing.written_fields(st):
        dev[fname].fill(0)
    stencil(**dev, **params, origin=origins, domain=domain)
    torch.cuda.synchronize()
    for fname in testing.written_fields(st):
        box = tuple(slice(o, o + d) for o, d in zip(origins[fname], domain))
        np.testing.assert_array_equal(dev[fname].get()[box], ref[fname][box])
