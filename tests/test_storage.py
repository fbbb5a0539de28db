"""Storage allocator contract (reference: tests/storage_tests/unit_tests/test_interface.py, test_layout.py)."""

import numpy as np
import pytest

from gt4py_b200 import storage


def test_layout_map_matches_gt_gpu_preset():
    # reference: storage/cartesian/layout.py:28-57 with base (2,1,0)
    assert storage.layout_map(("I", "J", "K")) == (2, 1, 0)
    assert storage.layout_map(("I", "J")) == (1, 0)
    assert storage.layout_map(("K",)) == (0,)
    assert storage.layout_map(("I", "J", "K", "0")) == (3, 2, 1, 0)
    assert storage.layout_map(("I", "J", "K", "0", "1")) == (4, 3, 2, 0, 1)


@pytest.mark.parametrize("shape,aligned", [((66, 66, 16), (1, 1, 0)), ((1028, 1028, 80), (2, 2, 0)), ((7, 5, 3), (0, 0, 0)), ((70, 50, 16), (3, 3, 0))])
def test_compute_layout_padding_and_alignment(shape, aligned):
    lm = storage.layout_map(("I", "J", "K"))
    strides, total, lead = storage.compute_layout(shape, lm, 8, 32, aligned)
    pitch = -(-shape[0] // 32) * 32
    assert strides == (1, pitch, pitch * shape[1])  # I unit stride, I padded to 32 elements
    assert total == pitch * shape[1] * shape[2]
    assert (lead + aligned[0]) % 32 == 0  # the aligned index sits on a 32-element boundary
    assert 0 <= lead < 32


def test_spec_errors_match_reference_types():
    with pytest.raises(TypeError):
        storage.normalize_storage_spec(None, 5, np.float64, None)
    with pytest.raises(ValueError):
        storage.normalize_storage_spec((0, 0), (4, 4, 4), np.float64, None)
    with pytest.raises(ValueError):
        storage.normalize_storage_spec(None, (4, 0, 4), np.float64, None)
    with pytest.raises(ValueError):
        storage.normalize_storage_spec(None, (4, 4, 4), np.float64, ("I", "J", "X"))
    ai, shape, dt, dims = storage.normalize_storage_spec(None, (4, 5), (np.float32, (3,)), ("I", "J"))
    assert shape == (4, 5, 3) and dims == ("I", "J", "0") and dt == np.float32 and ai == (0, 0, 0)


@pytest.mark.needs_gt4py
def test_strides_equal_reference_allocator():
    """Same strides / alignment as the reference's NDArrayBufferAllocator (allocators.py:187-273)
    driven through gt4py.storage for the registered b200 preset (CPU twin of the layout: numpy buffer)."""
    import gt4py_b200  # noqa: F401  (registers the preset)
    from gt4py.storage import allocators
    from gt4py.storage.cartesian import layout_registry, utils as gt_utils

    info = layout_registry.from_name("b200")
    rng = np.random.default_rng(0)
    for _ in range(20):
        shape = tuple(int(x) for x in rng.integers(1, 90, size=3))
        aligned = tuple(int(rng.integers(0, s)) for s in shape)
        for dtype in (np.float32, np.float64):
            lm = info["layout_map"](("I", "J", "K"))
            _, ref = gt_utils.allocate_cpu(shape, lm, np.dtype(dtype), info["alignment"] * np.dtype(dtype).itemsize, aligned)
            strides, _, lead = storage.compute_layout(shape, lm, np.dtype(dtype).itemsize, info["alignment"], aligned)
            assert tuple(s // ref.itemsize for s in ref.strides) == strides
            addr = ref.__array_interface__["data"][0] + sum(a * s for a, s in zip(aligned, ref.strides))
            assert addr % (32 * ref.itemsize) == 0
            assert (lead + aligned[0]) % 32 == 0


@pytest.mark.gpu
def test_device_array_surface():
    a = storage.zeros((70, 50, 16), np.float32, aligned_index=(3, 3, 0))
    assert a.shape == (70, 50, 16) and a.dtype == np.float32 and a.strides[0] == 4
    cai = a.__cuda_array_interface__
    assert (cai["data"][0] + 3 * a.strides[0] + 3 * a.strides[1]) % 128 == 0
    host = np.random.default_rng(1).random((70, 50, 16), dtype=np.float32)
    a[...] = host
    np.testing.assert_array_equal(a.get(), host)
    a[1:5, 2, :] = 7.0
    host[1:5, 2, :] = 7.0
    np.testing.assert_array_equal(np.asarray(a), host)
    b = storage.from_array(host, aligned_index=(3, 3, 0))
    v = b[3:, 3:, :]
    assert v.shape == (67, 47, 16) and v.data_ptr % 128 == 0
    t = b.transpose(2, 1, 0)
    assert t.shape == (16, 50, 70) and t.strides == tuple(reversed(b.strides))
    ones = storage.ones((4, 4), np.int32, dimensions=("I", "J"))
    assert int(ones.torch().sum()) == 16
    full = storage.full((3, 2, 2), 2.5, np.float64)
    assert float(full.torch().sum()) == 2.5 * 12
    assert storage.is_optimal_layout(b, ("I", "J", "K")) and not storage.is_optimal_layout(t, ("I", "J", "K"))


def test_device_array_behaves_like_an_ndarray_in_expressions(monkeypatch):
    """what user code and the reference's tests do with the cupy arrays of a GPU backend (comparisons, arithmetic,
    in-place updates, reductions, scalar conversion) on a b200 storage; evaluated by torch on the storage's device
    (host memory here: the allocator's device is stubbed)"""
    import torch

    from gt4py_b200 import storage

    monkeypatch.setattr(storage, "_device", lambda device=None: torch.device("cpu"))
    a_h = np.arange(24, dtype=np.float64).reshape(2, 3, 4)
    a = storage.from_array(a_h, aligned_index=(0, 0, 0))
    b = storage.full((2, 3, 4), 2.0, np.float64, aligned_index=(0, 0, 0))
    np.testing.assert_array_equal(((a + b) * 2 - 1).get(), (a_h + 2) * 2 - 1)
    np.testing.assert_array_equal((1.0 / (a + 1)).get(), 1.0 / (a_h + 1))
    np.testing.assert_array_equal((-a).get(), -a_h)
    np.testing.assert_array_equal(abs(a - 10).get(), abs(a_h - 10))
    assert (b == 2).all() and not (a == 2).all() and (a == 2).any() and (a[1:, :, 2] > 13).all()
    assert ((a >= 0) & (a < 24)).all() and not (~(a >= 0)).any()
    assert a.sum() == a_h.sum() and a.max() == 23 and a.min() == 0 and a.mean() == a_h.mean()
    np.testing.assert_array_equal(a.sum(axis=2).get(), a_h.sum(axis=2))
    a *= 2
    a += b
    a[0] -= 1
    np.testing.assert_array_equal(a.get(), np.concatenate([a_h[:1] * 2 + 1, a_h[1:] * 2 + 2]))
    assert float(a[0, 0, 0]) == 1.0 and int(a[1, 2, 3]) == 48 and bool(a[0, 0, 0] == 1) and a[0, 0, 1].item() == 3.0
    np.testing.assert_array_equal(a.reshape(6, 4).get(), a.get().reshape(6, 4))
    assert a.flatten().shape == (24,) and [float(x.sum()) for x in a] == [float(x.sum()) for x in a.get()]
    i32 = a.astype(np.int32)
    assert i32.dtype == np.int32 and i32.get()[1, 2, 3] == 48 and a.T.shape == (4, 3, 2)
    assert (storage.ones((3,), np.float32, aligned_index=(0,)) * np.float32(3)).dtype == np.float32
    assert np.all(b == 2) and np.sum(b) == 48.0 and not np.any(b > 2)  # NumPy's reductions dispatch to the methods
    r = np.float64(2.0) * b  # NumPy scalar on the left: still evaluated on the storage's device
    assert isinstance(r, storage.DeviceArray) and (r == 4).all()
    np.testing.assert_array_equal(np.asarray(b), np.full((2, 3, 4), 2.0))
    with pytest.raises(ValueError, match="ambiguous"):
        bool(a)
    with pytest.raises(TypeError, match="host NumPy"):
        a + a_h
    with pytest.raises(TypeError):
        hash(a)
