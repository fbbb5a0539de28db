"""IJ-decomposed multi-GPU execution: J-slab decomposition + NCCL SendRecv halo exchange.

The reference has no distributed path at all (no mpi4py / nccl / ghex call site in
gt4py.cartesian); domain decomposition is left to downstream users.  This module is the addition
north_star asks for: one process per GPU, the global IJK domain cut into P slabs along J, every
rank holding `nJ_local + 2h` rows of each exchanged field, and one `ncclSend/ncclRecv` group per
exchange on a dedicated stream so that the interior of the stencil can run concurrently.

With the backend's (2,1,0) layout a J-halo slab of width h is nK separate chunks of `h x pitch_I`
contiguous elements, so a slab is staged through a contiguous buffer by the launcher's strided copy
kernel (`b200_pack_2d`) before / after the NCCL call.

`torch.distributed` is used only as plumbing (rendezvous, barrier, broadcasting the NCCL unique id);
the data path is the C-ABI `b200_halo_exchange` (csrc/launcher.cu).  For CPU tests a `gloo`
transport moves the same slabs through host memory so the indexing logic is testable without GPUs.
"""

from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np


@dataclass(frozen=True)
class SlabDecomposition:
    """1-D decomposition of the J axis over `n_ranks` (non-periodic)."""

    n_ranks: int
    rank: int
    global_nj: int

    def __post_init__(self):
        if not (0 <= self.rank < self.n_ranks):
            raise ValueError(f"rank {self.rank} outside [0, {self.n_ranks})")
        if self.global_nj < self.n_ranks:
            raise ValueError("fewer J rows than ranks")

    def bounds(self, rank: Optional[int] = None) -> Tuple[int, int]:
        r = self.rank if rank is None else rank
        base, rem = divmod(self.global_nj, self.n_ranks)
        lo = r * base + min(r, rem)
        return lo, lo + base + (1 if r < rem else 0)

    @property
    def local_nj(self) -> int:
        lo, hi = self.bounds()
        return hi - lo

    @property
    def peer_lo(self) -> int:
        return self.rank - 1 if self.rank > 0 else -1

    @property
    def peer_hi(self) -> int:
        return self.rank + 1 if self.rank < self.n_ranks - 1 else -1

    def scatter(self, global_array: np.ndarray, halo: int, j_origin: int) -> np.ndarray:
        """Local slab (with `halo` rows on both sides) of a global array whose domain starts at
        row `j_origin` (which must be >= halo)."""
        lo, hi = self.bounds()
        return np.ascontiguousarray(global_array[:, j_origin + lo - halo : j_origin + hi + halo])


class HaloExchanger:
    """Exchanges the J-halos of a set of device fields with the two neighbouring ranks.

    fields: list of (DeviceArray | torch tensor, origin_j, halo_width); the local compute domain
    covers rows [origin_j, origin_j + local_nj) of each array.
    """

    def __init__(self, decomp: SlabDecomposition, local_nj: Optional[int] = None, *, transport: str = "nccl"):
        self.decomp = decomp
        self.local_nj = decomp.local_nj if local_nj is None else local_nj
        self.transport = transport
        self._comm = None
        self._plans: Dict[Any, Any] = {}
        self._stage_device = "cuda"
        self._stream = None
        if transport == "nccl":
            self._init_nccl()

    # -- NCCL bootstrap through torch.distributed (plumbing) ---------------------------------------
    def _init_nccl(self):
        import torch
        import torch.distributed as dist

        from . import runtime

        lib = runtime.load_library()
        uid = ctypes.create_string_buffer(128)
        if self.decomp.rank == 0:
            runtime.check(lib.b200_comm_unique_id(uid))
        payload = [bytes(uid.raw)]
        if self.decomp.n_ranks > 1:
            dist.broadcast_object_list(payload, src=0)
        comm = ctypes.c_void_p()
        idbuf = ctypes.create_string_buffer(payload[0], 128)
        runtime.check(lib.b200_comm_init(ctypes.byref(comm), idbuf, self.decomp.n_ranks, self.decomp.rank))
        self._comm = comm
        s = ctypes.c_void_p()
        # high priority: pack / NCCL / unpack get SM slots as soon as CTAs of a concurrently running stencil retire
        runtime.check(lib.b200_stream_create_priority(ctypes.byref(s), 1))
        self._stream = s
        self._lib = lib
        self._torch = torch

    @property
    def stream(self) -> int:
        return int(self._stream.value)

    # -- slab geometry -------------------------------------------------------------------------------
    @staticmethod
    def _slab_spec(view, j0: int, h: int):
        """A J-slab [j0, j0+h) of an (I, J, K) array as (ptr, rows, row_bytes, pitch_bytes).

        Needs I unit-stride and rows that are contiguous in J (stride_J == I pitch)."""
        ni, nj, nk = view.shape
        si, sj, sk = view.strides
        item = view.dtype.itemsize
        if si != 1:
            raise ValueError("halo exchange needs I-contiguous fields")
        if sk >= sj:  # layout (2,1,0): K outermost -> nK chunks of h*sj elements
            return view.ptr + j0 * sj * item, nk, h * sj * item, sk * item
        # layout (2,0,1): J outermost -> one contiguous chunk
        return view.ptr + j0 * sj * item, 1, h * sj * item, h * sj * item

    def _exchange_plan(self, views):
        """Everything one exchange of these buffers needs, prepared once per (pointer, geometry) set: the
        `b200_halo_t` array for the NCCL group, the argument tuples of the pack / unpack launches and the
        staging buffers.  The per-step cost of `exchange()` is then a handful of C calls."""
        from . import runtime

        key = tuple((v.ptr, v.shape, v.strides, oj, h) for v, oj, h in views)
        plan = self._plans.get(key)
        if plan is not None:
            return plan
        torch, d = self._torch, self.decomp
        halos = (runtime.B200Halo * max(1, len(views)))()
        packs: List[Tuple] = []
        unpacks: List[Tuple] = []
        keep = []
        for n, (view, oj, h) in enumerate(views):
            specs = {
                "send_lo": self._slab_spec(view, oj, h),
                "send_hi": self._slab_spec(view, oj + self.local_nj - h, h),
                "recv_lo": self._slab_spec(view, oj - h, h),
                "recv_hi": self._slab_spec(view, oj + self.local_nj, h),
            }
            nbytes = specs["send_lo"][1] * specs["send_lo"][2]
            stage = {k: torch.empty(nbytes, dtype=torch.uint8, device=self._stage_device) for k in specs}
            keep.append(stage)
            hd = halos[n]
            hd.bytes = nbytes
            for side, peer in (("lo", d.peer_lo), ("hi", d.peer_hi)):
                if peer < 0:
                    continue
                ptr, rows, row_bytes, pitch = specs[f"send_{side}"]
                if rows > 1:  # strided slab (layout (2,1,0): nK chunks): staged through a contiguous buffer
                    packs.append((stage[f"send_{side}"].data_ptr(), row_bytes, ptr, pitch, row_bytes, rows))
                    setattr(hd, f"send_{side}", stage[f"send_{side}"].data_ptr())
                    setattr(hd, f"recv_{side}", stage[f"recv_{side}"].data_ptr())
                    rptr, rrows, rrow_bytes, rpitch = specs[f"recv_{side}"]
                    unpacks.append((rptr, rpitch, stage[f"recv_{side}"].data_ptr(), rrow_bytes, rrow_bytes, rrows))
                else:  # contiguous slab: sent / received in place
                    setattr(hd, f"send_{side}", ptr)
                    setattr(hd, f"recv_{side}", specs[f"recv_{side}"][0])
        plan = (halos, packs, unpacks, keep)
        self._plans[key] = plan
        return plan

    def exchange(self, fields: Sequence[Tuple[Any, int, int]], *, stream: Optional[int] = None) -> int:
        """Enqueue pack -> NCCL send/recv -> unpack for all fields on `stream` (default: own stream).
        Returns the number of kernels launched (pack/unpack), NCCL calls not counted."""
        if self.transport != "nccl":
            raise RuntimeError("exchange() is the device path; use exchange_host() with the gloo transport")
        from . import runtime

        lib = self._lib
        st = self.stream if stream is None else stream
        halos, packs, unpacks, _keep = self._exchange_plan([(runtime.as_view(arr), oj, h) for arr, oj, h in fields])
        for args in packs:
            runtime.check(lib.b200_pack_2d(*args, st))
        runtime.check(lib.b200_halo_exchange(self._comm, halos, len(fields), self.decomp.peer_lo, self.decomp.peer_hi, st))
        for args in unpacks:
            runtime.check(lib.b200_pack_2d(*args, st))
        return len(packs) + len(unpacks)

    # -- host transport (gloo) for CPU tests of the indexing logic ----------------------------------
    def exchange_host(self, fields: Sequence[Tuple[np.ndarray, int, int]]) -> None:
        import torch
        import torch.distributed as dist

        d = self.decomp
        for arr, oj, h in fields:
            ops, recvs = [], []
            for peer, send_j, recv_j in (
                (d.peer_lo, oj, oj - h),
                (d.peer_hi, oj + self.local_nj - h, oj + self.local_nj),
            ):
                if peer < 0:
                    continue
                sbuf = torch.from_numpy(np.ascontiguousarray(arr[:, send_j : send_j + h]))
                rbuf = torch.empty_like(sbuf)
                ops += [dist.P2POp(dist.isend, sbuf, peer), dist.P2POp(dist.irecv, rbuf, peer)]
                recvs.append((recv_j, rbuf))
            if ops:
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
            for recv_j, rbuf in recvs:
                arr[:, recv_j : recv_j + h] = rbuf.numpy()

    def close(self):
        if self._comm is not None:
            self._lib.b200_comm_destroy(self._comm)
            self._comm = None
        if self._stream is not None:
            self._lib.b200_stream_destroy(self._stream)
            self._stream = None


class PeerHalo:
    """J-slab halo exchange over PEER MEMORY (NVLink / NVSwitch): no NCCL call, no staging, no pack / unpack kernels.

    Every exchanged field lives in symmetric memory (`torch.distributed._symmetric_memory`: the same allocation made on
    every rank and mapped into every process — plumbing).  Per step each rank runs ONE small kernel on its comm stream
    (`b200_halo_push`, csrc/launcher.cu) that stores its boundary rows straight into the neighbours' halo rows and then
    raises a flag in the neighbours' memory (system-scope release).  The consumer side is inside the stencil kernel:
    kernels generated with `halo_wait=True` run the whole slab in ONE launch, order the two boundary J tiles last and let
    them wait on the flags (device side, acquire) right before their first load — compute and exchange overlap tile by
    tile, without boundary-strip launches or host-side event chains.

    Ordering contract (one-step lag, two rotating buffer sets are enough): `push(fields)` must be enqueued after the
    kernel that produced the rows it sends (the caller makes the comm stream wait for that) and the step that consumes
    them must pass `wait_args()` of the same push."""

    def __init__(self, decomp: SlabDecomposition, local_nj: Optional[int] = None, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        from . import runtime

        self.decomp = decomp
        self.local_nj = decomp.local_nj if local_nj is None else local_nj
        self._torch, self._symm, self._lib = torch, symm, runtime.load_library()
        self._group = group if group is not None else dist.group.WORLD
        self._device = torch.device("cuda", torch.cuda.current_device())
        self._handles: List[Any] = []
        flags = symm.empty(8, dtype=torch.int64, device=self._device)  # [0] raised by rank-1, [1] by rank+1
        flags.zero_()
        torch.cuda.synchronize()
        hdl = symm.rendezvous(flags, self._group)
        self._flags, self._flag_ptrs = flags, [int(p) for p in hdl.buffer_ptrs]
        self._handles.append(hdl)
        hdl.barrier()
        self.epoch = 0
        s = ctypes.c_void_p()
        runtime.check(self._lib.b200_stream_create_priority(ctypes.byref(s), 1))
        self._stream = s
        self._plans: Dict[Any, Any] = {}

    @property
    def stream(self) -> int:
        return int(self._stream.value)

    def _raw_alloc(self, nbytes: int):
        return self._symm.empty(int(nbytes), dtype=self._torch.uint8, device=self._device)

    def _register(self, arr):
        """collective: rendezvous the allocation of `arr`; records where element [0,0,...] lives on every rank"""
        import torch.distributed as dist

        hdl = self._symm.rendezvous(arr._raw, self._group)
        self._handles.append(hdl)
        off = int(arr.data_ptr) - int(arr._raw.data_ptr())
        offs = [None] * self.decomp.n_ranks
        dist.all_gather_object(offs, off, group=self._group)
        arr._peer_ptrs = [int(p) + int(o) for p, o in zip(hdl.buffer_ptrs, offs)]
        return arr

    def empty(self, shape, dtype=np.float32, *, aligned_index=None, dimensions=None, fill=None):
        """Collective (every rank, same order, same arguments): a b200 storage in symmetric memory."""
        from . import storage as b2storage

        return self._register(b2storage.empty(shape, dtype, aligned_index=aligned_index, dimensions=dimensions, _fill=fill, raw_alloc=self._raw_alloc))

    def from_array(self, data, *, aligned_index=None, dimensions=None):
        from . import storage as b2storage

        return self._register(b2storage.from_array(data, aligned_index=aligned_index, dimensions=dimensions, raw_alloc=self._raw_alloc))

    def _plan(self, fields):
        from . import runtime

        key = tuple((int(a.data_ptr), oj, h) for a, oj, h in fields)
        plan = self._plans.get(key)
        if plan is not None:
            return plan
        d, nj = self.decomp, self.local_nj
        boxes = []
        for arr, oj, h in fields:
            if not hasattr(arr, "_peer_ptrs"):
                raise ValueError("PeerHalo: the field was not allocated with PeerHalo.empty / from_array")
            si, sj, sk = arr.element_strides
            item = arr.itemsize
            if si != 1 or arr.ndim != 3:
                raise ValueError("PeerHalo: fields must be 3-D and I-contiguous")
            nrow = ((h - 1) * sj + arr.shape[0]) * item  # h consecutive logical rows incl. the padding between them
            for peer, send_j, recv_j in ((d.peer_lo, oj, oj + nj), (d.peer_hi, oj + nj - h, oj - h)):
                if peer < 0:
                    continue
                b = runtime.B200Push()
                b.src = int(arr.data_ptr) + send_j * sj * item
                b.dst = arr._peer_ptrs[peer] + recv_j * sj * item
                b.row_bytes, b.rows, b.levels = nrow, 1, arr.shape[2]
                b.src_row_pitch = b.dst_row_pitch = sj * item
                b.src_level_pitch = b.dst_level_pitch = sk * item
                boxes.append(b)
        flags = []
        if d.peer_lo >= 0:
            flags.append(self._flag_ptrs[d.peer_lo] + 8)  # I am the lower neighbour's rank+1
        if d.peer_hi >= 0:
            flags.append(self._flag_ptrs[d.peer_hi] + 0)  # I am the upper neighbour's rank-1
        plan = ((runtime.B200Push * max(1, len(boxes)))(*boxes), len(boxes), (ctypes.c_void_p * max(1, len(flags)))(*flags), len(flags))
        self._plans[key] = plan
        return plan

    def push(self, fields: Sequence[Tuple[Any, int, int]], *, stream: Optional[int] = None) -> int:
        """Enqueue the push of the boundary rows of `fields` [(array, origin_j, halo width)] + the flag stores on `stream`
        (default: the exchanger's high-priority stream).  Returns the number of kernels launched."""
        from . import runtime

        boxes, nb, flags, nf = self._plan(fields)
        self.epoch += 1
        runtime.check(self._lib.b200_halo_push(boxes, nb, flags, nf, self.epoch, ctypes.c_void_p(self.stream if stream is None else stream)))
        return (1 if nb else 0) + (1 if nf else 0)

    def wait_args(self) -> Tuple[int, int, int]:
        """(flag_lo, flag_hi, epoch) for `FrozenStencil(..., halo_wait=...)` / b200_stencil_run_halo of the step that
        consumes the rows of the latest push."""
        me = self._flag_ptrs[self.decomp.rank]
        return (me if self.decomp.peer_lo >= 0 else 0, me + 8 if self.decomp.peer_hi >= 0 else 0, self.epoch)

    def wait(self, stream: int) -> int:
        """Consumer side for stencils without `halo_wait` kernels: a one-thread kernel on `stream` that returns when the
        rows of the latest push have landed (b200_halo_wait).  Returns the number of kernels launched."""
        from . import runtime

        lo, hi, epoch = self.wait_args()
        if not (lo or hi):
            return 0
        runtime.check(self._lib.b200_halo_wait(ctypes.c_void_p(lo or None), ctypes.c_void_p(hi or None), int(epoch), ctypes.c_void_p(stream)))
        return 1

    def barrier(self):
        self._handles[0].barrier()

    def close(self):
        if self._stream is not None:
            self._lib.b200_stream_destroy(self._stream)
            self._stream = None
