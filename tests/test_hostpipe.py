"""Host-resident call path (gt4py_b200/hostpipe.py): chunk planning + eligibility on CPU, the K-slab
composition through the emulator, and the real three-stream pipeline on the GPU (`-m gpu`)."""

import numpy as np
import pytest

from gt4py_b200 import hostpipe, testing
from oracle import numpy_oracle


def test_plan_chunks_cover_the_levels():
    for nk, n in ((80, 8), (80, 7), (5, 8), (1, 4), (160, 16)):
        ch = hostpipe.plan_chunks(nk, n)
        assert ch[0][0] == 0 and ch[-1][1] == nk and len(ch) == min(n, nk)
        assert all(a[1] == b[0] for a, b in zip(ch, ch[1:])) and all(hi > lo for lo, hi in ch)
        assert max(hi - lo for lo, hi in ch) - min(hi - lo for lo, hi in ch) <= 1


def test_slab_ranges_tile_the_flat_buffer():
    numel, offset, sk, ko, nk = 7 + 10 * 96, 7, 96, 1, 8  # 10 levels of 96 elements, lead 7, K origin 1
    ch = hostpipe.plan_chunks(nk, 3)
    r = [hostpipe.slab_range(numel, offset, sk, ko, k0, k1, c == 0, c == len(ch) - 1) for c, (k0, k1) in enumerate(ch)]
    assert r[0][0] == 0 and r[-1][1] == numel
    assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
    # level k of the domain lives at [offset + (ko + k) * sk, +sk): inside its chunk's range
    for c, (k0, k1) in enumerate(ch):
        for k in range(k0, k1):
            assert r[c][0] <= offset + (ko + k) * sk and offset + (ko + k + 1) * sk <= r[c][1]


@pytest.mark.parametrize(
    "name,ok",
    [("hdiff_f32", True), ("upwind5_f32", True), ("laplacian_f64", True), ("tridiagonal_f64", False), ("k_intervals_f64", False),
     ("tmp_koffset_f64", False), ("kiter_f64", False), ("varoff_f64", False), ("lowdim_f64", False)],
)  # fmt: skip
def test_pipeline_eligibility(name, ok):
    why = hostpipe.pipeline_supported(testing.load_ir(name, "staged"))
    assert (why is None) == ok, why


def test_k_slabs_compose_to_the_full_call_emulated():
    """what HostPipeline launches: the stencil on K-sliced views, chunk by chunk == one full call"""
    from emu.emu import EmuStencil

    st = testing.load_ir("hdiff_f32", "staged")
    fields, params, origins, domain = testing.make_case_data(st, "hdiff_f32", domain=(37, 9, 7), seed=3)
    ref = {k: v.copy() for k, v in fields.items()}
    numpy_oracle.run(st, ref, params, domain, origins)
    es = EmuStencil(st, {}, name="hdiff_f32.staged")
    for k0, k1 in hostpipe.plan_chunks(domain[2], 3):
        views = {n: a[:, :, origins[n][2] + k0 : origins[n][2] + k1] for n, a in fields.items()}
        org = {n: (o[0], o[1], 0) for n, o in origins.items()}
        es.run(views, params, (domain[0], domain[1], k1 - k0), org)
    np.testing.assert_array_equal(fields["out_field"], ref["out_field"])


@pytest.mark.gpu
@pytest.mark.parametrize("name,chunks", [("hdiff_f32", 4), ("hdiff_f32", 1), ("upwind5_f32", 3)])
def test_host_pipeline_matches_oracle(name, chunks):
    import torch

    from gt4py_b200 import storage
    from gt4py_b200.stencil import B200Stencil

    st = testing.load_ir(name, "staged")
    fields, params, origins, domain = testing.make_case_data(st, name, domain=(150, 40, 9), seed=8)
    ref = {k: v.copy() for k, v in fields.items()}
    numpy_oracle.run(st, ref, params, domain, origins)
    stencil = B200Stencil(st, {"device_sync": False})
    # device staging: inputs start zeroed (they must arrive through the pipeline's H2D copies); outputs start from
    # the host content, because the D2H copy returns the WHOLE pitched buffer and the stencil leaves halo cells as is
    written = testing.written_fields(st)
    dev = {k: (storage.from_array(v, aligned_index=origins[k]) if k in written else storage.zeros(v.shape, v.dtype, aligned_index=origins[k]))
           for k, v in fields.items()}  # fmt: skip
    host = {k: hostpipe.PinnedMirror(dev[k], v) for k, v in fields.items()}
    pipe = hostpipe.HostPipeline(stencil, dev, origin=origins, domain=domain, n_chunks=chunks)
    for _ in range(2):  # twice: the device staging buffers are reused across calls
        assert pipe(**host, **params) >= len(pipe.chunks)
    torch.cuda.synchronize()
    for fname in testing.written_fields(st):
        np.testing.assert_array_equal(host[fname].array, ref[fname], err_msg=f"{name}:{fname}")
    with pytest.raises(ValueError, match="K slabs are not independent"):
        t = testing.load_ir("tridiagonal_f64", "default")
        hostpipe.HostPipeline(B200Stencil(t), {}, origin={}, domain=(4, 4, 4))
