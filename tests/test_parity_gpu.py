"""-m gpu: the CUDA path (through the C-ABI launcher) against the oracle and the reference goldens."""

import numpy as np
import pytest

from gt4py_b200 import storage, testing

from parity_util import run_case

pytestmark = pytest.mark.gpu
CASES = testing.list_cases()


@pytest.mark.parametrize("variant", ["default", "staged"])
@pytest.mark.parametrize("name", CASES)
def test_fixture_parity_point_generator(name, variant):
    """Every fixture, baseline generator, vs oracle AND vs the stored outputs of the reference."""
    run_case(name, variant, {"strategy": "point"}, check_golden=True)


@pytest.mark.parametrize("name", CASES)
def test_fixture_parity_auto_strategy(name):
    """What `backend="b200"` picks by default (streaming kernels where applicable)."""
    run_case(name, "staged", {"strategy": "auto"}, check_golden=True)
    run_case(name, "default", {"strategy": "auto"}, seed=1, check_golden=True)


@pytest.mark.parametrize("name", ["hdiff_f32", "upwind5_f32", "laplacian_f64", "two_stage_par_f32", "fw_pgrad_f32", "stage_halo_f32"])
@pytest.mark.parametrize("domain", [(1, 1, 1), (3, 2, 1), (127, 5, 2), (129, 33, 3), (260, 70, 4)])
def test_ragged_domains(name, domain):
    for variant in ("default", "staged"):
        run_case(name, variant, None, domain=domain, seed=2)


@pytest.mark.parametrize("name", ["tridiagonal_f64", "vadv_f64", "fw_wsolve_f32", "fwd_scan_f64", "col_mask_f64",
                                  "col_chain_f64", "col_backward_f64", "col_multiwrite_f32"])  # fmt: skip
@pytest.mark.parametrize("domain", [(1, 1, 4), (70, 3, 9), (130, 37, 33)])
def test_column_solvers_domains(name, domain):
    """register k-cache column kernels (default), without prefetch, and the baseline column kernel"""
    run_case(name, "default", None, domain=domain, seed=4)
    run_case(name, "staged", {"seq_prefetch": False}, domain=domain, seed=5)
    run_case(name, "default", {"seq_cache": False}, domain=domain, seed=6)
    run_case(name, "default", {"fuse_columns": True, "seq_prefetch": 2}, domain=domain, seed=7)


def test_c_order_torch_tensors_any_stride():
    """Plain C-ordered (K-contiguous) torch tensors must be accepted (reference:
    tests/cartesian_tests/integration_tests/feature_tests/test_field_layouts.py:31-47); layout only warns."""
    import torch

    with pytest.warns(UserWarning, match="layout"):
        run_case("hdiff_f32", "staged", None, to_device=lambda n, a: torch.from_numpy(a).cuda())
    with pytest.warns(UserWarning, match="layout"):
        run_case("tridiagonal_f64", "default", None, to_device=lambda n, a: torch.from_numpy(a).cuda())


def test_gt_dims_permutation():
    """__gt_dims__ = ('K','J','I') arrays are transposed to IJK before the call
    (reference: stencil_object.py:78-92, test_call_interface.py)."""
    import torch

    class Permuted:
        def __init__(self, arr):
            self.t = torch.from_numpy(np.ascontiguousarray(arr.transpose(2, 1, 0))).cuda()
            self.__gt_dims__ = ("K", "J", "I")
            self.__cuda_array_interface__ = self.t.__cuda_array_interface__

        def cpu(self):
            return self.t.permute(2, 1, 0).cpu()

    run_case("laplacian_f64", "default", None, domain=(20, 13, 5), to_device=lambda n, a: Permuted(a))


def test_hdiff_full_size_properties():
    """BASELINE.json configs[1] at full size (1024x1024x80 fp32): size-independent properties.
    * a constant field is a fixed point of horizontal diffusion (lap == 0 -> out == in), exactly
    * default and staged lowerings and both generators agree bit for bit on random data
    * a sub-box of the full run equals the oracle run on that sub-box alone (translation invariance)."""
    from gt4py_b200.stencil import B200Stencil

    st_s = testing.load_ir("hdiff_f32", "staged")
    st_d = testing.load_ir("hdiff_f32", "default")
    n, nk, h = 1024, 80, 2
    if __import__("os").environ.get("B200_EMULATE_DEVICE") == "1":  # logic check of this test on the fake device
        n, nk = 96, 4
    rng = np.random.default_rng(5)
    shape = (n + 2 * h, n + 2 * h, nk)
    origin = {k: (h, h, 0) for k in ("in_field", "out_field", "coeff")}
    inp = storage.from_array(rng.random(shape, dtype=np.float32), aligned_index=(h, h, 0))
    coeff = storage.from_array(rng.random(shape, dtype=np.float32) * np.float32(0.1), aligned_index=(h, h, 0))
    outs = []
    for st, opts in ((st_s, {"strategy": "auto"}), (st_d, {"strategy": "point"}), (st_s, {"strategy": "point"})):
        out = storage.zeros(shape, np.float32, aligned_index=(h, h, 0))
        B200Stencil(st, opts)(inp, out, coeff, origin=origin, domain=(n, n, nk))
        outs.append(out.torch().clone())
    import torch

    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    # oracle on a corner sub-box that includes the halo
    from oracle import numpy_oracle

    sub = 48
    f = {
        "in_field": inp.get()[: sub + 2 * h, : sub + 2 * h, :4].copy(),
        "coeff": coeff.get()[: sub + 2 * h, : sub + 2 * h, :4].copy(),
        "out_field": np.zeros((sub + 2 * h, sub + 2 * h, 4), np.float32),
    }
    numpy_oracle.run(st_d, f, {}, (sub, sub, 4), origin)
    got = outs[0][h : h + sub, h : h + sub, :4].cpu().numpy()
    np.testing.assert_array_equal(got, f["out_field"][h : h + sub, h : h + sub, :])
    # fixed point
    const = storage.full(shape, 3.25, np.float32, aligned_index=(h, h, 0))
    out = storage.zeros(shape, np.float32, aligned_index=(h, h, 0))
    B200Stencil(st_s, None)(const, out, coeff, origin=origin, domain=(n, n, nk))
    assert torch.equal(out.torch()[h:-h, h:-h, :], const.torch()[h:-h, h:-h, :])


def test_hdiff_full_size_whole_domain_against_the_oracle():
    """BASELINE.json configs[1] at FULL size, every cell: the oracle (NumPy restatement of the reference numpy backend,
    ~10 s for 84 M cells) against the variants that are actually timed — the backend's default (lazy per-pitch
    specialisation: interior loop + compile-time pitch + uniform task index), the bulk-async (TMA) variants bench.py's
    autotune selects among, and the `halo_wait` kernel of the multi-GPU schedule — all bit for bit."""
    import torch

    from gt4py_b200.stencil import B200Stencil
    from oracle import numpy_oracle

    st = testing.load_ir("hdiff_f32", "staged")
    n, nk, h = 1024, 80, 2
    if __import__("os").environ.get("B200_EMULATE_DEVICE") == "1":  # logic check of this test on the fake device
        n, nk = 130, 3
    rng = np.random.default_rng(17)
    shape = (n + 2 * h, n + 2 * h, nk)
    origin = {k: (h, h, 0) for k in ("in_field", "out_field", "coeff")}
    host = {"in_field": rng.random(shape, dtype=np.float32), "coeff": rng.random(shape, dtype=np.float32) * np.float32(0.1),
            "out_field": np.zeros(shape, np.float32)}  # fmt: skip
    inp = storage.from_array(host["in_field"], aligned_index=(h, h, 0))
    coeff = storage.from_array(host["coeff"], aligned_index=(h, h, 0))
    numpy_oracle.run(st, host, {}, (n, n, nk), origin)
    want = torch.from_numpy(host["out_field"]).to(inp._base.device)
    pitch = inp.element_strides[1]
    P = {"interior_loop": True, "static_pitch": pitch}
    variants = [{"specialize": "lazy"}, dict(P), {**P, "tma": 3, "tile_j": 32, "prefetch": 1}, {**P, "tma": 3, "tile_j": 32, "prefetch": 1, "tma_mode": "bulk"},
                {**P, "tma": 4, "prefetch": 1}, {**P, "halo_wait": True}, {**P, "uniform_task": False}]  # fmt: skip
    for opts in variants:
        out = storage.zeros(shape, np.float32, aligned_index=(h, h, 0))
        stencil = B200Stencil(st, {"strategy": "auto", **opts})
        stencil(inp, out, coeff, origin=origin, domain=(n, n, nk))
        got = out.torch()
        assert torch.equal(got, want), f"{opts}: {(got != want).sum().item()} cells differ from the oracle"
        if opts.get("specialize") == "lazy":  # the default really ran the specialised kernels
            assert stencil.compiled._special and next(iter(stencil.compiled._special.values())).options.get("static_pitch") == pitch


@pytest.mark.parametrize("name,variant", [("hdiff_f32", "staged"), ("hdiff_f32", "default"), ("fw_pgrad_f32", "staged"), ("upwind5_f32", "staged")])
def test_subbox_launches_compose_to_the_full_domain(name, variant):
    """`b200_stencil_run(..., subbox)` (used to overlap the halo exchange with interior compute):
    interior + boundary strips launched separately must give exactly the full-domain result."""
    from gt4py_b200.stencil import B200Stencil
    from oracle import numpy_oracle

    st = testing.load_ir(name, variant)
    fields, params, origins, domain = testing.make_case_data(st, name, domain=(150, 200, 4), seed=9)
    ref = {k: v.copy() for k, v in fields.items()}
    numpy_oracle.run(st, ref, params, domain, origins)
    for strategy in ("auto", "point"):
        dev = {k: storage.from_array(v, aligned_index=origins[k]) for k, v in fields.items()}
        fr = B200Stencil(st, {"strategy": strategy}).freeze(origin=origins, domain=domain)
        ni, nj, _ = domain
        for box in ((0, ni, 64, nj - 64), (0, ni, 0, 64), (0, ni, nj - 64, nj)):
            fr(**dev, **params, subbox=box)
        for fname in testing.written_fields(st):
            np.testing.assert_array_equal(dev[fname].get(), ref[fname], err_msg=f"{name}.{variant}/{strategy}:{fname}")
        # I-direction split with unaligned cuts as well
        dev = {k: storage.from_array(v, aligned_index=origins[k]) for k, v in fields.items()}
        for box in ((0, 37, 0, nj), (37, 101, 0, 3), (37, 101, 3, nj), (101, ni, 0, nj)):
            fr(**dev, **params, subbox=box)
        for fname in testing.written_fields(st):
            np.testing.assert_array_equal(dev[fname].get(), ref[fname], err_msg=f"{name}.{variant}/{strategy}:{fname} (I split)")


def _smooth_case(ni, nj, nk, seed):
    """SURVEY §8d second input family: the smooth analytic field of the reference's Burgers demo
    (5 + 8 (2 + cos(pi (x + 1.5 y)) + sin(2 pi (x + 1.5 y))) / 4) with a constant coefficient 0.025: the flux
    limiter then sits on exact sign changes and plateaus (products that are exactly zero), not on noise"""
    st = testing.load_ir("hdiff_f32", "staged")
    fields, params, origins, domain = testing.make_case_data(st, "hdiff_f32", domain=(ni, nj, nk), seed=seed)
    shp = fields["in_field"].shape
    x = np.linspace(0.0, 1.0, shp[0], dtype=np.float64)[:, None, None]
    y = np.linspace(0.0, 1.0, shp[1], dtype=np.float64)[None, :, None]
    z = 1.0 + 0.1 * np.arange(shp[2], dtype=np.float64)[None, None, :]
    fields["in_field"] = (z * (5.0 + 8.0 * (2.0 + np.cos(np.pi * (x + 1.5 * y)) + np.sin(2.0 * np.pi * (x + 1.5 * y))) / 4.0)).astype(np.float32)
    fields["coeff"] = np.full(fields["coeff"].shape, 0.025, np.float32)
    return st, fields, params, origins, domain


@pytest.mark.parametrize("options", [None, {"strategy": "point"}, {"interior_loop": True, "specialize": "lazy"}])
def test_hdiff_smooth_analytic_field(options):
    from gt4py_b200.stencil import B200Stencil
    from oracle import numpy_oracle

    st, fields, params, origins, domain = _smooth_case(260, 130, 3, seed=1)
    ref = {k: v.copy() for k, v in fields.items()}
    numpy_oracle.run(st, ref, params, domain, origins)
    dev = {k: storage.from_array(v, aligned_index=origins[k]) for k, v in fields.items()}
    B200Stencil(st, options)(**dev, origin=origins, domain=domain)
    np.testing.assert_array_equal(dev["out_field"].get(), ref["out_field"])
    box = tuple(slice(o, o + d) for o, d in zip(origins["out_field"], domain))
    assert np.abs(ref["out_field"][box] - ref["in_field"][tuple(slice(o, o + d) for o, d in zip(origins["in_field"], domain))]).max() > 0


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.bool_, np.int64])
@pytest.mark.parametrize("shape,dims,ai", [((70, 45, 9), None, (2, 3, 0)), ((33, 17), ("I", "J"), (1, 1)), ((19,), ("K",), (0,)),
                                            ((64, 64, 16), None, (0, 0, 0)), ((5, 3, 130), None, (1, 0, 0))])
def test_from_array_upload_goes_through_the_relayout_kernel(dtype, shape, dims, ai):
    """storage.from_array: one contiguous H2D + the launcher's tiled re-layout kernel (b200_relayout) == the host data,
    for C-ordered, Fortran-ordered and sliced sources"""
    from gt4py_b200 import storage

    rng = np.random.default_rng(7)
    src = (rng.random(shape) * 100).astype(dtype) if dtype != np.bool_ else rng.random(shape) > 0.5
    for host in (src, np.asfortranarray(src), np.ascontiguousarray(np.flip(src, 0))[::-1]):
        dev = storage.from_array(host, aligned_index=ai, dimensions=dims)
        assert dev.dtype == np.dtype(dtype) and dev.shape == shape
        np.testing.assert_array_equal(dev.get(), src)
