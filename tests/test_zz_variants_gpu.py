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


ROUND2_VARIANTS = [
    ("upwind5_f32", "staged", {"period": 8, "div_slow": "inline"}, (262, 150, 2)),
    ("upwind5_f32", "staged", {"div_inv": False, "period": 4}, (262, 150, 2)),
    ("upwind5_f32", "staged", {"interior_loop": True, "static_pitch": 288, "warps": 2}, (262, 150, 2)),
    ("fw_pgrad_f32", "staged", {"k_order": False}, (200, 70, 6)),
    ("fw_pgrad_f32", "staged", {"k_order": True, "prefetch": 1, "thin_tile_j": 16}, (200, 70, 6)),
    ("fw_pgrad_f32", "staged", {"interior_loop": True, "static_pitch": 224, "tma": 3, "tile_j": 32, "prefetch": 0}, (200, 70, 6)),
    ("fw_div_f32", "staged", {"interior_loop": True, "static_pitch": 224, "tma": 3, "tile_j": 16, "prefetch": 1, "tma_mode": "bulk"}, (200, 70, 6)),
    ("fw_wsolve_f32", "default", {"col_smem": True, "seq_prefetch": 4}, (150, 40, 30)),
    ("fw_wsolve_f32", "default", {"col_smem": True}, (40, 9, 130)),  # taller than the shared-memory budget: global-scratch path
    ("fw_wsolve_f32", "default", {"seq_rotate": False}, (150, 40, 30)),
    ("tridiagonal_f64", "default", {"seq_rotate": True, "seq_prefetch": 4}, (150, 40, 30)),
    ("tridiagonal_f64", "default", {"fuse_columns": True, "seq_prefetch": 5, "col_hints": True}, (150, 40, 30)),
    ("vadv_f64", "default", {"col_smem": True, "col_smem_kb": 100, "seq_prefetch": 2}, (70, 33, 20)),
    ("vadv_f64", "default", {"seq_prefetch": 2, "col_hints": True}, (70, 33, 20)),
]


@pytest.mark.parametrize("name,variant,opts,domain", ROUND2_VARIANTS)
def test_round2_generator_options_match_the_oracle_on_the_device(name, variant, opts, domain):
    """unroll factor / division / task order / look-ahead / shared-memory variants of round 2: same bits as the oracle"""
    from parity_util import run_case

    run_case(name, variant, {"device_sync": True, **opts}, domain=domain, seed=41)


def test_division_fallback_on_the_device_gives_the_reference_bits():
    """zeros (both signs), infinities, subnormals and huge dividends leave the guarded exponent range of the
    hoisted-reciprocal division: the warp redoes the quotient group with IEEE divisions (b200::div_ieee) — same bits as NumPy"""
    import torch

    from gt4py_b200 import storage
    from gt4py_b200.stencil import B200Stencil
    from oracle import numpy_oracle

    st = testing.load_ir("upwind5_f32", "staged")
    fields, params, origins, domain = testing.make_case_data(st, "upwind5_f32", domain=(262, 150, 2), seed=43)
    fields["u"][5:140, 7:60, :] = 0.0
    fields["u"][50, 30, 1] = -0.0
    fields["v"][20:30, 3:9, 0] = np.inf
    fields["v"][60, 8, 1] = 1e-44
    fields["u"][200:220, 100:120, :] *= np.float32(1e30)
    ref = {k: v.copy() for k, v in fields.items()}
    with np.errstate(all="ignore"):
        numpy_oracle.run(st, ref, params, domain, origins)
    for opts in ({}, {"interior_loop": True, "static_pitch": 288}, {"div_slow": "inline", "period": 4}):
        dev = {k: storage.from_array(v, aligned_index=origins[k]) for k, v in fields.items()}
        B200Stencil(st, {"device_sync": True, **opts})(**dev, **params, origin=origins, domain=domain)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(dev["out"].get().view(np.uint32), ref["out"].view(np.uint32), err_msg=str(opts))
