"""The REAL plug-in path end to end — `@gtscript.stencil(backend="b200")`, `gt4py.storage.*(backend="b200")`,
`StencilObject.__call__` -> generated `run()` -> `run_compiled` -> launcher call — compared with the reference
`numpy` backend run in the same process on the same inputs.

Every test runs twice:
  * `[fake]`  CPU suite: on the fake device (tests/emu/fake_device.py: host memory, launches through the kernel emulator)
  * `[cuda]`  `-m gpu`: on the B200, with the reference package from `baseline/_ref/` (tools/install_reference.sh)
Reference call path being replaced: stencil_object.py:531-612, backend/templates/stencil_module.py.in:91-169,
storage/cartesian/interface.py:264-327."""

import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.needs_gt4py


@pytest.fixture(scope="module", params=[pytest.param("fake"), pytest.param("cuda", marks=pytest.mark.gpu)])
def fake_device(request):
    try:
        import torch
    except ImportError:
        pytest.skip("torch missing")
    if request.param == "cuda":
        if not torch.cuda.is_available():
            from emu import fake_device as fd

            if not fd.installed():
                pytest.skip("no CUDA device")
        yield "cuda"
        return
    if torch.cuda.is_available():
        pytest.skip("real device present")
    from emu import fake_device as fd

    if fd.installed():  # B200_EMULATE_DEVICE=1 session: already on the fake device
        yield "fake"
        return
    mp = pytest.MonkeyPatch()
    fd.install(mp)
    yield "fake"
    mp.undo()


def _stencils(backend):
    from gt4py.cartesian import gtscript
    from gt4py.cartesian.gtscript import BACKWARD, FORWARD, PARALLEL, Field, computation, interval

    F = Field[np.float64]

    @gtscript.stencil(backend=backend, rebuild=True)
    def smooth(u: F, out: F, *, alpha: np.float64):
        with computation(PARALLEL), interval(...):
            lap = 4.0 * u[0, 0, 0] - (u[1, 0, 0] + u[-1, 0, 0] + u[0, 1, 0] + u[0, -1, 0])
            out = u - alpha * (4.0 * lap[0, 0, 0] - (lap[1, 0, 0] + lap[-1, 0, 0] + lap[0, 1, 0] + lap[0, -1, 0]))

    @gtscript.stencil(backend=backend, rebuild=True)
    def cumsum_and_back(a: F, s: F):
        with computation(FORWARD):
            with interval(0, 1):
                s = a
            with interval(1, None):
                s = s[0, 0, -1] + a
        with computation(BACKWARD), interval(0, -1):
            s = s + 0.5 * s[0, 0, 1]

    return smooth, cumsum_and_back


def test_b200_stencils_match_the_numpy_backend_through_the_gt4py_call_path(fake_device):
    warnings.filterwarnings("ignore")
    import gt4py.storage as gt_storage

    import gt4py_b200  # noqa: F401  (registers backend="b200")

    rng = np.random.default_rng(4)
    shape, origin, domain = (27, 19, 6), (2, 2, 0), (23, 15, 6)
    u_h, a_h = rng.random(shape), rng.random(shape)
    results = {}
    for backend in ("numpy", "b200"):
        smooth, scan = _stencils(backend)
        u = gt_storage.from_array(u_h, backend=backend, aligned_index=origin)
        out = gt_storage.zeros(shape, np.float64, backend=backend, aligned_index=origin)
        a = gt_storage.from_array(a_h, backend=backend, aligned_index=(0, 0, 0))
        s = gt_storage.zeros(shape, np.float64, backend=backend, aligned_index=(0, 0, 0))
        info = {}
        smooth(u, out, alpha=np.float64(0.05), origin=origin, domain=domain, exec_info=info)
        scan(a, s)  # origin / domain inferred from the arguments
        assert "run_end_time" in info or "call_end_time" in info
        results[backend] = (np.asarray(out) if backend == "numpy" else out.get(), np.asarray(s) if backend == "numpy" else s.get())
        if backend == "b200":
            assert smooth.backend == "b200" and info.get("b200_kernel_launches", 0) >= 1
            assert u.strides[0] == 8 and u.strides[2] > u.strides[1] > u.strides[0]  # the backend's (2,1,0) pitched layout
    np.testing.assert_array_equal(results["b200"][0], results["numpy"][0])
    np.testing.assert_array_equal(results["b200"][1], results["numpy"][1])
    assert np.abs(results["b200"][0]).sum() > 0 and np.abs(results["b200"][1]).sum() > 0


def test_lower_dimensional_data_dim_and_masked_fields_through_the_plugin(fake_device):
    """IJ-only / K-only fields, a data-dimension field, integer fields with an `if`, storage constructors and host
    copies of b200 storages — all through gt4py's own interfaces, against the numpy backend"""
    warnings.filterwarnings("ignore")
    import gt4py.storage as gt_storage
    from gt4py.cartesian import gtscript
    from gt4py.cartesian.gtscript import IJ, PARALLEL, Field, K, computation, interval

    import gt4py_b200  # noqa: F401

    def build(backend):
        @gtscript.stencil(backend=backend, rebuild=True)
        def mixed(a: Field[np.float64], sfc: Field[IJ, np.float64], prof: Field[K, np.float64],
                  vec: Field[(np.float64, (3,))], flag: Field[np.int32], out: Field[np.float64], *, w: np.float64):  # fmt: skip
            with computation(PARALLEL), interval(...):
                tmp = a * prof + sfc + vec[0, 0, 0][1] * w
                if flag > 0:
                    out = tmp + a[1, 0, 0]
                else:
                    out = tmp - a[0, -1, 0]

        return mixed

    rng = np.random.default_rng(9)
    ni, nj, nk = 18, 11, 5
    host = {
        "a": rng.random((ni + 2, nj + 2, nk)), "sfc": rng.random((ni + 2, nj + 2)), "prof": rng.random((nk,)),
        "vec": rng.random((ni + 2, nj + 2, nk, 3)), "flag": rng.integers(-2, 3, (ni + 2, nj + 2, nk)).astype(np.int32),
    }  # fmt: skip
    dims = {"a": "IJK", "sfc": "IJ", "prof": "K", "vec": ["I", "J", "K", "0"], "flag": "IJK"}
    outs = {}
    for backend in ("numpy", "b200"):
        st = build(backend)
        args = {}
        for n, h in host.items():
            d = list(dims[n])
            ai = tuple(1 if x in "IJ" else 0 for x in d)
            args[n] = gt_storage.from_array(h, h.dtype, backend=backend, aligned_index=ai, dimensions=d)
        out = gt_storage.full((ni + 2, nj + 2, nk), -7.0, np.float64, backend=backend, aligned_index=(1, 1, 0))
        st(**args, out=out, w=np.float64(1.5), origin={"_all_": (1, 1, 0), "prof": (0,), "sfc": (1, 1)}, domain=(ni, nj, nk))
        outs[backend] = np.asarray(out)  # (__array__ of the b200 storage = device -> host copy)
        if backend == "b200":
            ones = gt_storage.ones((4, 5, 6), np.float32, backend="b200", aligned_index=(0, 0, 0))
            assert np.asarray(ones).sum() == 120 and ones.dtype == np.float32
            sl = ones[1:3, :, 2]  # slicing keeps a device view
            assert tuple(sl.shape) == (2, 5)
    np.testing.assert_array_equal(outs["b200"], outs["numpy"])
    assert (outs["b200"][1:-1, 1:-1] != -7.0).all() and (outs["b200"][0] == -7.0).all()


def test_frozen_stencil_backend_options_and_rebuild_from_cache(fake_device):
    """gt4py's `freeze()`, backend options (`device_sync`, the b200 code-generation options) and a second build of
    the same definition (served from .gt_cache with the persisted cubin + plan) through the real call path"""
    warnings.filterwarnings("ignore")
    import gt4py.storage as gt_storage
    from gt4py.cartesian import gtscript
    from gt4py.cartesian.gtscript import PARALLEL, Field, computation, interval

    import gt4py_b200  # noqa: F401
    from gt4py_b200 import backend as b2backend

    F = Field[np.float32]

    def definition(u: F, c: F, out: F):
        with computation(PARALLEL), interval(...):
            lap = 4.0 * u[0, 0, 0] - (u[1, 0, 0] + u[-1, 0, 0] + u[0, 1, 0] + u[0, -1, 0])
            out = u - c * (lap[1, 0, 0] - lap[-1, 0, 0] + lap[0, 1, 0] - lap[0, -1, 0])

    rng = np.random.default_rng(12)
    shape, origin, domain = (40, 22, 4), (2, 2, 0), (36, 18, 4)
    u_h, c_h = rng.random(shape, dtype=np.float32), rng.random(shape, dtype=np.float32)
    ref = gtscript.stencil(backend="numpy", definition=definition, name="frz", literal_float_precision=32)
    mk = lambda a, b: gt_storage.from_array(a, np.float32, backend=b, aligned_index=origin)  # noqa: E731
    o_ref = gt_storage.zeros(shape, np.float32, backend="numpy", aligned_index=origin)
    ref(mk(u_h, "numpy"), mk(c_h, "numpy"), o_ref, origin=origin, domain=domain)
    for n, opts in enumerate(({"device_sync": False}, {"device_sync": False, "interior_loop": True, "specialize": "lazy"}, {})):
        st = gtscript.stencil(backend="b200", definition=definition, name="frz", literal_float_precision=32, rebuild=(n < 2), **opts)
        assert st.backend == "b200"
        u, c = mk(u_h, "b200"), mk(c_h, "b200")
        out = gt_storage.zeros(shape, np.float32, backend="b200", aligned_index=origin)
        frozen = st.freeze(origin={"u": origin, "c": origin, "out": origin}, domain=domain)
        frozen(u=u, c=c, out=out)
        np.testing.assert_array_equal(out.get(), np.asarray(o_ref), err_msg=str(opts))
        out2 = gt_storage.zeros(shape, np.float32, backend="b200", aligned_index=origin)
        st(u, c, out2, origin=origin, domain=domain, validate_args=False)
        np.testing.assert_array_equal(out2.get(), np.asarray(o_ref))
    assert len(b2backend._COMPILED) >= 1


def test_fusing_and_graphing_gt4py_stencil_objects(fake_device):
    """two applications of a `backend="b200"` StencilObject as ONE fused stencil (fuse.fuse_stencils) and as one
    captured graph (StencilGraph) == two calls of the numpy-backend stencil"""
    warnings.filterwarnings("ignore")
    import gt4py.storage as gt_storage
    from gt4py.cartesian import gtscript
    from gt4py.cartesian.gtscript import PARALLEL, Field, computation, interval

    import gt4py_b200  # noqa: F401
    from gt4py_b200 import fuse
    from gt4py_b200.graph import StencilGraph

    F = Field[np.float64]

    def definition(u: F, out: F):
        with computation(PARALLEL), interval(...):
            out = 0.5 * u + 0.125 * (u[1, 0, 0] + u[-1, 0, 0] + u[0, 1, 0] + u[0, -1, 0])

    rng = np.random.default_rng(21)
    shape, dom = (30, 26, 3), (26, 22, 3)
    u_h = rng.random(shape)
    ref = gtscript.stencil(backend="numpy", definition=definition, name="avg")
    mid, out = np.zeros(shape), np.zeros(shape)
    ref(u_h, mid, origin=(1, 1, 0), domain=(dom[0] + 2, dom[1] + 2, dom[2]))  # step 1 wherever step 2 reads it
    ref(mid, out, origin=(2, 2, 0), domain=dom)

    st = gtscript.stencil(backend="b200", definition=definition, name="avg", device_sync=False)
    two = fuse.fuse_stencils("avg_x2", [(st, {"out": "mid"}), (st, {"u": "mid"})], intermediates=["mid"], options={"device_sync": True})
    assert two.field_info["u"].boundary[:2] == ((2, 2), (2, 2))
    u = gt_storage.from_array(u_h, backend="b200", aligned_index=(2, 2, 0))
    o1 = gt_storage.zeros(shape, np.float64, backend="b200", aligned_index=(2, 2, 0))
    two(u=u, out=o1, origin=(2, 2, 0), domain=dom)
    np.testing.assert_array_equal(o1.get(), out)

    # the same two steps as a captured sequence of the gt4py object's calls
    m2 = gt_storage.zeros(shape, np.float64, backend="b200", aligned_index=(2, 2, 0))
    o2 = gt_storage.zeros(shape, np.float64, backend="b200", aligned_index=(2, 2, 0))
    st(u, m2, origin=(1, 1, 0), domain=(dom[0] + 2, dom[1] + 2, dom[2]))  # warm-up (also allocates nothing here)
    m2[...] = 0.0
    g = StencilGraph()
    with g:
        st(u, m2, origin=(1, 1, 0), domain=(dom[0] + 2, dom[1] + 2, dom[2]))
        st(m2, o2, origin=(2, 2, 0), domain=dom)
    assert float(np.abs(o2.get()).sum()) == 0.0 and g.num_nodes == 2
    g.launch()
    np.testing.assert_array_equal(o2.get(), out)
    g.close()


# ---------------------------------------------------------------------------------------------------------------------
# every fixture stencil of tools/stencil_defs.py (benchmark stencils + one OIR feature each) through the plug-in call
# path: built twice from the SAME GTScript definition (backend="numpy" = the oracle north_star names, backend="b200"),
# gt4py.storage allocation per backend, StencilObject.__call__ with the same seeded inputs, bit-for-bit comparison
# ---------------------------------------------------------------------------------------------------------------------
#: run on the fake device too (the emulator compiles each kernel with g++: a few seconds per case)
_FAKE_CASES = ("hdiff_f32", "tridiagonal_f64", "regions_f64", "datadims_f64")
#: fixtures that call libm transcendentals: NumPy's and CUDA's libm agree to 1e-11 relative, not bit for bit
_RTOL = {"math_f64": 1e-11, "rounding_f64": 1e-11}


def _registry_cases():
    import pathlib
    import sys

    tools = str(pathlib.Path(__file__).resolve().parent.parent / "tools")
    if tools not in sys.path:
        sys.path.insert(0, tools)
    try:
        import stencil_defs
    except Exception:  # gt4py not importable: the module-level mark skips everything anyway
        return list(_FAKE_CASES)
    return list(stencil_defs.REGISTRY)


def _dims_of(decl):
    return [a for a, on in zip("IJK", decl["dims"]) if on] + [str(n) for n in range(len(decl["data_dims"]))]


@pytest.mark.parametrize("case_name", _registry_cases())
def test_fixture_stencils_through_the_plugin_equal_the_numpy_backend(fake_device, case_name):
    warnings.filterwarnings("ignore")
    if fake_device == "fake" and case_name not in _FAKE_CASES and not __import__("os").environ.get("B200_PLUGIN_ALL"):
        pytest.skip("fake device: representative subset only (every case runs on the GPU)")
    import gt4py.storage as gt_storage
    import stencil_defs
    from gt4py.cartesian import gtscript

    import gt4py_b200  # noqa: F401
    from gt4py_b200 import testing

    case = stencil_defs.REGISTRY[case_name]
    ir = testing.load_ir(case_name, "default")
    decls = {p["name"]: p for p in ir["params"] if p["t"] == "field"}
    fields, params, origins, domain = testing.make_case_data(ir, case_name, seed=5)
    outs = {}
    for backend in ("numpy", "b200"):
        st = gtscript.stencil(backend=backend, definition=case["definition"], externals=case["externals"] or {},
                              name=f"{case_name}_{backend}_plug", **case["build"])  # fmt: skip
        assert st.backend == backend
        args = {}
        for n, h in fields.items():
            if h is None:
                args[n] = None
                continue
            # (a field only accessed above / below a K offset has a negative boundary, hence a negative origin: legal as a
            #  call origin, not as a storage alignment hint)
            args[n] = gt_storage.from_array(h, h.dtype, backend=backend, aligned_index=tuple(max(0, int(o)) for o in origins[n][: h.ndim]),
                                            dimensions=_dims_of(decls[n]))  # fmt: skip
        info = {}
        st(**args, **params, origin=origins, domain=domain, exec_info=info)
        outs[backend] = {n: np.asarray(args[n]) for n in testing.written_fields(ir)}
        if backend == "b200":
            assert info.get("b200_kernel_launches", 0) >= 1
            assert "run_device_time" in info and info["run_device_time"] > 0.0  # device-side counterpart of run_cpp_*_time
    for n in testing.written_fields(ir):
        if case_name in _RTOL:
            np.testing.assert_allclose(outs["b200"][n], outs["numpy"][n], rtol=_RTOL[case_name], atol=0, equal_nan=True)
        else:
            np.testing.assert_array_equal(outs["b200"][n], outs["numpy"][n], err_msg=f"{case_name}:{n}")
        assert not np.array_equal(outs["b200"][n], fields[n]) or fields[n].size == 0


def test_host_arrays_through_the_plugin_call(fake_device):
    """`StencilObject.__call__` with HOST arrays (an extension: the reference's GPU backends refuse CPU arrays): plain
    NumPy arguments take a pageable copy + re-layout per field; `gt4py_b200.storage.host_*` arrays (pinned, same pitched
    layout as the device storages) are moved with contiguous DMAs, K-slab pipelined with the kernels when the levels are
    independent.  Results land in the caller's host arrays, bit-equal to the numpy backend's."""
    warnings.filterwarnings("ignore")
    import stencil_defs
    from gt4py.cartesian import gtscript

    import gt4py_b200  # noqa: F401
    from gt4py_b200 import storage as b2storage, testing

    for case_name, expect_path in (("hdiff_f32", "pipeline"), ("tridiagonal_f64", "serial")):
        case = stencil_defs.REGISTRY[case_name]
        ir = testing.load_ir(case_name, "default")
        fields, params, origins, domain = testing.make_case_data(ir, case_name, domain=(70, 33, 12), seed=8)
        ref = gtscript.stencil(backend="numpy", definition=case["definition"], externals=case["externals"] or {}, name=f"{case_name}_h_ref", **case["build"])
        st = gtscript.stencil(backend="b200", definition=case["definition"], externals=case["externals"] or {}, name=f"{case_name}_h_b200", **case["build"])
        want = {n: v.copy() for n, v in fields.items()}
        ref(**want, **params, origin=origins, domain=domain)
        # (a) plain NumPy arrays, C order
        got = {n: v.copy() for n, v in fields.items()}
        info = {}
        st(**got, **params, origin=origins, domain=domain, exec_info=info)
        assert info["b200_host_path"] == "serial" and info["b200_kernel_launches"] >= 1
        for n in testing.written_fields(ir):
            np.testing.assert_array_equal(got[n], want[n], err_msg=f"{case_name}:{n} (numpy arguments)")
        # (b) pinned host storages in the backend's layout
        pinned = {n: b2storage.host_from_array(v, aligned_index=origins[n]) for n, v in fields.items()}
        assert all(isinstance(a, np.ndarray) and a.strides[0] == a.itemsize for a in pinned.values())
        info = {}
        st(**pinned, **params, origin=origins, domain=domain, exec_info=info)
        assert info["b200_host_path"] == expect_path, info
        for n in testing.written_fields(ir):
            np.testing.assert_array_equal(np.asarray(pinned[n]), want[n], err_msg=f"{case_name}:{n} (pinned host storages)")
        if all(fi is None or fi["access"] != "READ_WRITE" for fi in ir["field_info"].values()):  # (idempotent stencils only)
            halo_probe = pinned[testing.written_fields(ir)[0]]
            halo_probe[0, 0, 0] = -123.0  # outside the compute domain: must survive the call (only the domain travels back)
            st(**pinned, **params, origin=origins, domain=domain)  # second call: cached staging / pipeline
            assert halo_probe[0, 0, 0] == -123.0
            halo_probe[0, 0, 0] = want[testing.written_fields(ir)[0]][0, 0, 0]
            for n in testing.written_fields(ir):
                np.testing.assert_array_equal(np.asarray(pinned[n]), want[n])
