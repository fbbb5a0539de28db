"""The DEVICE path of the halo exchange (`HaloExchanger.exchange`: pack -> NCCL send/recv group -> unpack,
cached exchange plan) executed on host memory: the three C-ABI entry points it calls are stood in by a host
implementation (strided copy; send/recv rendezvous between threads, one thread per rank).  Checks the slab
geometry, the staging and the peer / side bookkeeping of the real code path for 2 and 3 ranks (a middle rank
exchanges on both sides), for the backend's (2,1,0) layout (strided slabs, staged) and for a J-outermost layout (2,0,1) (contiguous slabs,
sent in place)."""

import ctypes
import threading

import numpy as np
import pytest
import torch

from gt4py_b200 import distributed, runtime, storage


class _HostFabric:
    """b200_pack_2d / b200_halo_exchange on host pointers; ranks rendezvous at a barrier."""

    def __init__(self, n_ranks):
        self.n = n_ranks
        self.barrier = threading.Barrier(n_ranks)
        self.posted = {}
        self.calls = {r: [] for r in range(n_ranks)}

    def lib(self, rank):
        fabric = self

        class Lib:
            def b200_pack_2d(self, dst, dst_pitch, src, src_pitch, row_bytes, rows, stream):
                fabric.calls[rank].append(("pack", rows, row_bytes))
                for r in range(rows):
                    ctypes.memmove(dst + r * dst_pitch, src + r * src_pitch, row_bytes)
                return 0

            def b200_halo_exchange(self, comm, halos, n_halos, peer_lo, peer_hi, stream):
                fabric.calls[rank].append(("exchange", n_halos, peer_lo, peer_hi))
                fabric.posted[rank] = [(h.send_lo, h.recv_lo, h.send_hi, h.recv_hi, h.bytes) for h in halos[:n_halos]]
                fabric.barrier.wait()
                for n in range(n_halos):
                    _slo, rlo, _shi, rhi, nbytes = fabric.posted[rank][n]
                    if peer_lo >= 0:  # my low halo <- the high boundary rows of the rank below
                        ctypes.memmove(rlo, fabric.posted[peer_lo][n][2], nbytes)
                    if peer_hi >= 0:
                        ctypes.memmove(rhi, fabric.posted[peer_hi][n][0], nbytes)
                fabric.barrier.wait()
                return 0

        return Lib()


class _HostExchanger(distributed.HaloExchanger):
    def __init__(self, decomp, local_nj, lib):
        self._host_lib = lib
        super().__init__(decomp, local_nj, transport="nccl")

    def _init_nccl(self):  # no NCCL / device here: the launcher calls go to the host fabric
        self._lib, self._torch, self._comm, self._stream = self._host_lib, torch, ctypes.c_void_p(1), ctypes.c_void_p(7)
        self._stage_device = "cpu"


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("layout", ["b200", "j_outer"])
def test_device_exchange_path_on_host_memory(monkeypatch, world, layout):
    monkeypatch.setattr(storage, "_device", lambda device=None: torch.device("cpu"))
    real_as_view = runtime.as_view

    def as_view(obj):  # host tensors are refused by the product (no CPU path): accepted here for the test only
        if isinstance(obj, torch.Tensor):
            return runtime.ArrayView(obj.data_ptr(), obj.shape, obj.stride(), np.dtype(str(obj.dtype).replace("torch.", "")), obj)
        return real_as_view(obj)

    monkeypatch.setattr(runtime, "as_view", as_view)
    ni, nj, nk, h = 21, 12, 4, 3
    rng = np.random.default_rng(3)
    glob = rng.random((ni, nj * world + 2 * h, nk), dtype=np.float32)
    fabric = _HostFabric(world)
    local, errors = {}, []

    def rank_main(rank):
        try:
            dec = distributed.SlabDecomposition(world, rank, nj * world)
            ex = _HostExchanger(dec, nj, fabric.lib(rank))
            mine = dec.scatter(glob, h, h).copy()
            lo, hi = dec.bounds()
            if rank > 0:
                mine[:, :h] = -1.0  # halos start wrong: they must arrive through the exchange
            if rank < world - 1:
                mine[:, -h:] = -1.0
            def j_outer(a):  # I unit-stride, then K, J outermost: strides (1, nK*nI, nI)
                return torch.from_numpy(np.ascontiguousarray(a.transpose(1, 2, 0))).permute(2, 0, 1)

            arr = storage.from_array(mine, aligned_index=(0, h, 0)) if layout == "b200" else j_outer(mine)
            other = storage.from_array(mine * 2, aligned_index=(0, h, 0)) if layout == "b200" else j_outer(mine * 2)
            for _ in range(2):  # second call: the cached plan
                n = ex.exchange([(arr, h, h), (other, h, h)])
            sides = (rank > 0) + (rank < world - 1)
            assert n == (2 * 2 * sides if layout == "b200" else 0)  # (2,1,0): pack + unpack per side and field; J outermost: in place
            local[rank] = (arr.get() if layout == "b200" else arr.numpy(), other.get() if layout == "b200" else other.numpy())
        except Exception as exc:  # surface assertion errors of the worker threads
            errors.append((rank, exc))
            fabric.barrier.abort()

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=60)
    assert not errors, errors
    for rank in range(world):
        dec = distributed.SlabDecomposition(world, rank, nj * world)
        expect = dec.scatter(glob, h, h)
        np.testing.assert_array_equal(local[rank][0], expect)
        np.testing.assert_array_equal(local[rank][1], expect * 2)
    assert [c[0] for c in fabric.calls[0]].count("exchange") == 2
