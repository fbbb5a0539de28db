"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/gt4py_b200.h declares; entry points fail loudly (never fall back) without a device."""

import ctypes
import pathlib
import re

import pytest

from gt4py_b200 import jit, runtime

ROOT = pathlib.Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "gt4py_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    jit.build_launcher()
    lib = ctypes.CDLL(str(runtime.LIB_PATH))
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in gt4py_b200.h but not exported"
    assert set(syms) == set(runtime.EXPORTED_SYMBOLS)
    assert runtime.load_library().b200_abi_version() == 1


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("device present")
    from gt4py_b200 import storage, testing
    from gt4py_b200.stencil import B200Stencil

    with pytest.raises(RuntimeError):
        storage.zeros((4, 4, 4))
    st = B200Stencil(testing.load_ir("copy_f64"))  # code generation + nvcc work without a GPU
    view = runtime.ArrayView(0x1000, (4, 4, 4), (1, 4, 16), "float64")
    with pytest.raises(runtime.B200Error):
        st.compiled.run_views({"a": view, "b": view}, b"", (2, 2, 2), {"a": (0, 0, 0), "b": (0, 0, 0)}, stream=0)


def test_bad_plan_is_rejected():
    lib = runtime.load_library()
    h = ctypes.c_void_p()
    rc = lib.b200_stencil_load(b"xx", 2, b"not a plan", ctypes.byref(h))
    assert rc < 0 and b"plan" in lib.b200_last_error()


def test_lazy_pitch_specialisation_selects_one_variant_per_pitch(monkeypatch):
    """specialize="lazy": descriptors with a common unit-stride row pitch get the static-pitch +
    interior-loop variant of the streaming kernels (compiled once per pitch); anything else keeps
    the generic kernels.  Code generation + nvcc only — nothing is launched here."""
    from gt4py_b200 import testing

    st = testing.load_ir("hdiff_f32", "staged")
    cs = runtime.CompiledStencil(st, {"specialize": "lazy"})
    org = {n: (2, 2, 0) for n in ("in_field", "out_field", "coeff")}

    def descs(pitch, si=1):
        v = runtime.ArrayView(0x10000, (36, 20, 4), (si, pitch, pitch * 20), "float32")
        return cs.make_field_descs({"in_field": v, "out_field": v, "coeff": v}, org)

    a = cs.specialized_for(descs(64))
    assert a is not cs and a.options["static_pitch"] == 64 and a.options["interior_loop"]
    assert "constexpr long long SJ = 64;" in a.source and "winterior" in a.source
    assert cs.specialized_for(descs(64)) is a  # cached
    assert cs.specialized_for(descs(96)).options["static_pitch"] == 96
    assert cs.specialized_for(descs(64, si=2)) is cs  # not I-contiguous: generic kernels
    assert runtime.CompiledStencil(st, {"specialize": "off"}).specialized_for(descs(64)).options.get("static_pitch") is None
    # the backend's default is "lazy" (the suites pin GT4PY_B200_SPECIALIZE=off, tests/conftest.py)
    monkeypatch.delenv("GT4PY_B200_SPECIALIZE", raising=False)
    dflt = runtime.CompiledStencil(st, {}).specialized_for(descs(64))
    assert dflt.options["static_pitch"] == 64 and dflt.options["interior_loop"] is True
    col = runtime.CompiledStencil(testing.load_ir("tridiagonal_f64"), {"specialize": "lazy"})
    assert not col._special and all(k["kind"] != "stream" for k in col.plan["kernels"])


def test_autotune_control_flow_with_a_stubbed_device(monkeypatch):
    """bench.py runs B200Stencil.autotune on its main path: exercise its control flow here (candidate
    resolution incl. static_pitch="auto", de-duplication, parity check against the first candidate,
    rejection of a candidate that writes something else, selection) with launches and CUDA events stubbed."""
    import numpy as np
    import torch

    from gt4py_b200 import storage, testing
    from gt4py_b200.stencil import B200Stencil

    st = testing.load_ir("hdiff_f32", "staged")
    shape, org = (40, 24, 3), (2, 2, 0)

    def dev(fill):
        es, total, lead = storage.compute_layout(shape, storage.layout_map(("I", "J", "K")), 4, 32, org)
        return storage.DeviceArray(torch.full((total + lead,), fill, dtype=torch.float32), lead, shape, es, np.float32)

    fields = {"in_field": dev(1.0), "coeff": dev(0.5), "out_field": dev(0.0)}
    calls = []

    def fake_run(self, descs, scalars, domain, *, stream=None, subbox=None, halo_wait=None):
        calls.append(dict(self.options))
        bad = self.options.get("tile_j") == 32  # this variant "computes" something else
        fields["out_field"].torch().fill_(7.0 if not bad else 8.0)
        return 1

    class FakeEvent:
        t = 0.0

        def __init__(self, enable_timing=False):
            pass

        def record(self):
            FakeEvent.t += 1.0
            self.at = FakeEvent.t

        def synchronize(self):
            pass

        def elapsed_time(self, other):
            # pretend the static-pitch interior variant is the fastest
            last = calls[-1]
            return 1.0 if (last.get("interior_loop") is True and last.get("static_pitch") and len(last) == 3) else 2.0 + len(last)

    monkeypatch.setattr(runtime.CompiledStencil, "run_descs", fake_run)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    stencil = B200Stencil(st, {"device_sync": False})
    cands = [{}, {"static_pitch": "auto"}, {"interior_loop": True, "static_pitch": "auto"}, {"tile_j": 32}, {"tile_j": 64}]
    tuned = stencil.autotune(fields, {}, domain=(36, 20, 3), origin={n: org for n in fields}, candidates=cands, iters=2)
    assert stencil.tune_rejected == [{"tile_j": 32}]
    assert tuned[0][0] == {"interior_loop": True, "static_pitch": 64} and stencil.backend_options["static_pitch"] == 64
    assert [c for c, _ in tuned].count({"tile_j": 64}) == 0  # same source as the default: de-duplicated
    assert len(tuned) == 3
    json_ok = __import__("json").dumps({"autotune": tuned, "options": stencil.backend_options})
    assert "static_pitch" in json_ok


def test_stencil_graph_control_flow_with_a_stubbed_launcher(monkeypatch):
    """StencilGraph: capture on a private stream when the current stream is the (uncapturable) legacy
    default stream, route launches that pass no stream to it for the duration of the block, restore
    afterwards, close the capture when the body raises."""
    import ctypes

    from gt4py_b200 import graph

    log = []

    class FakeLib:
        def b200_stream_create(self, ref):
            ctypes.cast(ref, ctypes.POINTER(ctypes.c_void_p))[0] = 0x5000
            return 0

        def b200_graph_begin(self, s):
            log.append(("begin", s.value))
            return 0

        def b200_graph_end(self, s, ref):
            log.append(("end", s.value))
            ctypes.cast(ref, ctypes.POINTER(ctypes.c_void_p))[0] = 0x7000
            return 0

        def b200_graph_num_nodes(self, h):
            return 3

        def b200_graph_launch(self, h, s):
            log.append(("launch", h.value, s.value or 0))
            return 0

        def b200_graph_destroy(self, h):
            log.append(("destroy", h.value))
            return 0

        def b200_stream_destroy(self, s):
            return 0

    monkeypatch.setattr(runtime, "load_library", lambda *a, **k: FakeLib())
    monkeypatch.setattr(runtime, "_stream_override", type("L", (), {})())
    import torch

    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a: type("S", (), {"cuda_stream": 0})())
    g = graph.StencilGraph()
    with g:
        assert runtime.current_stream_handle() == 0x5000 == g.stream  # launches are redirected
    assert runtime.current_stream_handle() == 0 and log == [("begin", 0x5000), ("end", 0x5000)]
    assert g.num_nodes == 3
    g.launch()
    assert log[-1] == ("launch", 0x7000, 0)
    with pytest.raises(runtime.B200Error, match="already holds"):
        with g:
            pass
    g2 = graph.StencilGraph(stream=0x9000)  # an explicit capturable stream is used as is
    with pytest.raises(KeyError):
        with g2:
            assert g2.stream == 0x9000
            raise KeyError("body failed")
    assert runtime.current_stream_handle() == 0 and log[-1] == ("destroy", 0x7000)
    with pytest.raises(runtime.B200Error, match="nothing captured"):
        g2.launch()


def test_frozen_stencil_descriptor_cache(monkeypatch):
    """FrozenStencil keeps prepared descriptors per set of argument objects (re-validated against their pointers
    and shapes): alternating buffer sets hit the cache, a re-pointed tensor does not, scalars are repacked on change."""
    import ctypes

    import numpy as np
    import torch

    from gt4py_b200 import storage, testing
    from gt4py_b200.stencil import B200Stencil

    seen = []

    class Lib:
        def b200_stencil_run(self, handle, descs, nfields, scalars, nbytes, dom, sb, stream):
            seen.append(([int(descs[n].data or 0) for n in range(nfields)], bytes(scalars), tuple(dom), None if sb is None else tuple(sb)))
            return 1

        def __getattr__(self, name):
            return lambda *a: 0

    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(storage, "_device", lambda device=None: torch.device("cpu"))
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a: type("S", (), {"cuda_stream": 0})())
    monkeypatch.setattr(runtime, "load_library", lambda *a, **k: Lib())
    monkeypatch.setattr(runtime.CompiledStencil, "handle", property(lambda self: ctypes.c_void_p(1)))
    st = testing.load_ir("upwind5_f32", "staged")
    fields, params, origins, domain = testing.make_case_data(st, "upwind5_f32", domain=(40, 20, 2), seed=0)
    a = {k: storage.from_array(v, aligned_index=origins[k]) for k, v in fields.items()}
    b = {k: storage.from_array(v, aligned_index=origins[k]) for k, v in fields.items()}
    fr = B200Stencil(st, {"device_sync": False}).freeze(origin=origins, domain=domain)
    names = [p["name"] for p in st["params"] if p["t"] == "field"]
    base = lambda s: [s[n].data_ptr + sum(o * e for o, e in zip((0, 0, 0), s[n].element_strides)) for n in names]  # noqa: E731
    for s in (a, b, a, b):
        fr(**s, **params)
    assert [c[0] for c in seen] == [base(a), base(b), base(a), base(b)] and len(fr._cache) == 2
    assert seen[0][1] == seen[1][1] and seen[0][2] == domain and seen[0][3] is None
    p2 = dict(params)
    first = next(iter(p2))
    p2[first] = type(p2[first])(0.125)
    fr(**a, **p2, subbox=(0, 40, 4, 16))
    assert seen[-1][1] != seen[0][1] and seen[-1][3] == (0, 40, 4, 16)
    # same objects, one of them re-pointed in place (torch set_): the cached descriptors must not be reused
    t = {k: v.torch().clone() for k, v in a.items()}
    monkeypatch.setattr(runtime, "as_view", lambda o: runtime.ArrayView(o.data_ptr(), o.shape, o.stride(), np.dtype("float32"), o)
                        if isinstance(o, torch.Tensor) else runtime.ArrayView(o.data_ptr, o.shape, o.element_strides, o.dtype, o))  # fmt: skip
    fr(**t, **params)
    before = seen[-1][0]
    t[names[0]].set_(t[names[0]].clone())
    fr(**t, **params)
    assert seen[-1][0][0] != before[0] and seen[-1][0][1:] == before[1:]


def test_launcher_accepts_every_generated_plan():
    """The C launcher parses the launch plan before it touches CUDA: without a device a well-formed plan must fail
    with NO_DEVICE (or CUDA), never with INVALID — checked for every fixture x lowering x generator and for fused IRs."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("device present")
    from gt4py_b200 import codegen, fuse, testing

    lib = runtime.load_library()
    hd = testing.load_ir("hdiff_f32", "staged")
    irs = [(testing.load_ir(n, v), o) for n in testing.list_cases() for v in ("default", "staged") for o in ({"strategy": "auto"}, {"strategy": "point"})]
    irs.append((fuse.repeat(hd, 2, carry=("in_field", "out_field")), {}))
    irs.append((testing.load_ir("tridiagonal_f64", "default"), {"fuse_columns": True, "seq_prefetch": 4, "seq_smem_pad": 110 * 1024}))
    bad = []
    for st, opts in irs:
        _src, plan = codegen.generate(st, opts)
        h = ctypes.c_void_p()
        rc = lib.b200_stencil_load(b"\\x7fELF-not-a-real-cubin", 22, codegen.plan_to_text(plan).encode(), ctypes.byref(h))
        msg = lib.b200_last_error().decode(errors="replace")
        if rc >= 0 or "plan" in msg:
            bad.append((st["name"], opts, rc, msg))
    assert not bad, bad[:3]
