"""CPU emulation of the b200 launch path for tests: compiles the GENERATED CUDA source with g++
(-DB200_HOST_EMU, tests/emu/cuda_shim.h) and executes the launch plan on host arrays.

Test infrastructure only.  It re-implements, in Python, what csrc/launcher.cu does for one call
(argument block, scratch layout of temporaries, grid geometry, step sequencing) so that the code
generators can be checked against the oracle without a GPU.
"""

from __future__ import annotations

import ctypes
import hashlib
import mmap
import os
import pathlib
import struct
import subprocess
import tempfile

import numpy as np

from gt4py_b200 import codegen, ir as b2ir

HERE = pathlib.Path(__file__).resolve().parent
CSRC = HERE.parent.parent / "gt4py_b200" / "csrc"
BUILD = pathlib.Path(tempfile.gettempdir()) / "gt4py_b200_emu"


def _runtime_object() -> pathlib.Path:
    key = hashlib.sha256((HERE / "cuda_shim.h").read_bytes() + (HERE / "emu_runtime.cpp").read_bytes()).hexdigest()[:16]
    obj = BUILD / f"emu_runtime_{key}.o"
    if not obj.exists():
        tmp = f"{obj}.{os.getpid()}.tmp"
        cmd = ["g++", "-std=c++20", "-O1", "-pthread", "-fPIC", "-c", "-DB200_HOST_EMU", "-I", str(HERE),
               str(HERE / "emu_runtime.cpp"), "-o", tmp]  # fmt: skip
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError("emulator runtime build failed:\n" + proc.stderr[-3000:])
        os.replace(tmp, obj)
    return obj


def _align_up(v, a):
    return (v + a - 1) // a * a


class EmuStencil:
    def __init__(self, stencil_ir, options=None, name="emu"):
        self.ir = stencil_ir
        self.source, self.plan = codegen.generate(stencil_ir, dict(options or {}))
        BUILD.mkdir(parents=True, exist_ok=True)
        tramp = "\n".join(
            f'extern "C" void emu_call_{k["name"]}(const void* blob) {{ {k["name"]}(*reinterpret_cast<const Args*>(blob)); }}'
            for k in self.plan["kernels"]
        )
        text = self.source + "\n" + tramp + "\n"
        key = hashlib.sha256(text.encode() + (HERE / "cuda_shim.h").read_bytes() + (HERE / "emu_runtime.cpp").read_bytes() + (CSRC / "b200_device.cuh").read_bytes()).hexdigest()[:20]
        so = BUILD / f"{codegen._cname(name)}_{key}.so"
        if not so.exists():
            src = BUILD / f"{codegen._cname(name)}_{key}.cpp"
            src.write_text(text)
            # the lockstep runtime (heavy <thread>/<barrier> headers) is compiled once; kernels at -O0:
            # compile time dominates these tests, the domains are tiny
            cmd = ["g++", "-std=c++20", "-O0", "-pthread", "-shared", "-fPIC", "-DB200_HOST_EMU", "-ffp-contract=off",
                   "-fsanitize=alignment", "-fno-sanitize-recover=alignment",  # misaligned vector access = device trap
                   "-I", str(HERE), "-I", str(CSRC), "-x", "c++", str(src), "-x", "none", str(_runtime_object()), "-o", str(so) + ".tmp"]
            proc = subprocess.run(cmd, capture_output=True, text=True)
            if proc.returncode != 0:
                raise RuntimeError("emulator build failed:\n" + proc.stderr[-3000:])
            os.replace(str(so) + ".tmp", so)
        self.lib = ctypes.CDLL(str(so))
        self.lib.emu_launch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint), ctypes.POINTER(ctypes.c_uint), ctypes.c_int]
        self.lib.emu_trace_read.restype = ctypes.c_longlong
        self.lib.emu_trace_read.argtypes = [ctypes.c_int, ctypes.c_int]
        self.launches = 0

    def trace(self, reset=True):
        """B200_TRACE counters: {0: steady-loop trips (pure), 1: steady-loop trips (edge/fastall),
        2: general march steps, 3: steady-loop trips (interior), 4: interior prologue/epilogue steps, 6: quotient groups redone by IEEE division (per lane)} of the streaming kernels since the last reset."""
        return {slot: int(self.lib.emu_trace_read(slot, int(reset))) for slot in range(8)}

    # ---- mirror of launcher.cu ------------------------------------------------------------------
    def _temp_layout(self, f, dom):
        e = f["extent"] or [[0, 0], [0, 0]]
        (ei0, ei1), (ej0, ej1) = e
        dims = f["dims"]
        nj = dom[1] + (ej1 - ej0) if dims[1] else 1
        nk = dom[2] if dims[2] else 1
        lead = _align_up(-ei0, 32) if dims[0] else 0
        pitch = _align_up(lead + dom[0] + ei1, 32) if dims[0] else 1
        nd = int(np.prod(f["data_dims"])) if f["data_dims"] else 1
        vol = pitch * nj * nk
        s = [1 if dims[0] else 0, pitch if dims[1] else 0, pitch * nj if dims[2] else 0, 0, 0, 0, 0]
        acc = vol
        for d in range(len(f["data_dims"]) - 1, -1, -1):
            s[3 + d] = acc
            acc *= f["data_dims"][d]
        origin = [lead if dims[0] else 0, -ej0 if dims[1] else 0, 0]
        return s, origin, vol * nd, nk

    def run(self, fields, params, domain, origins, subbox=None, layout=None, guard=None, halo_wait=(0, 0, 0)):
        """fields: name -> numpy arrays in IJK[+data] axis order (modified in place).

        layout="b200": stage every field in the backend's storage layout first (I unit-stride, rows
        padded to 32 elements, origin aligned — gt4py_b200.storage.compute_layout), which is what
        selects the 16-byte vector path / the steady-state loop of the streaming kernels; results
        are copied back.  guard="end"|"start": the staged buffers end (start) flush against an
        inaccessible page, so any access outside the allocation the GPU launcher would be given
        faults here instead of passing silently."""
        if layout == "b200":
            staged, keep = {}, []
            for name, arr in fields.items():
                if arr is None:
                    staged[name] = None
                    continue
                staged[name] = _stage_b200(arr, origins[name], guard, keep)
            self.run(staged, params, domain, origins, subbox=subbox, halo_wait=halo_wait)
            for name, arr in fields.items():
                if arr is not None:
                    arr[...] = staged[name]
            del staged, keep
            return
        plan = self.plan
        dom = [int(d) for d in domain]
        nf = len(plan["fields"])
        keep = []
        fa = []
        geo = {}  # field index -> (address of array element [0,0,0], shape3, strides, origin3): what launcher.cu encodes tensor maps from
        for fidx, f in enumerate(plan["fields"]):
            item = f["itemsize"]
            if f["kind"] == "dead":
                fa.append((0, [0] * 7, 0, 0, 0))
                continue
            if f["kind"] == "temp":
                s, org, nelem, nk = self._temp_layout(f, dom)
                buf = np.zeros(nelem + 64, dtype=np.dtype("bool" if f["dtype"] == "bool" else f["dtype"]))
                keep.append(buf)
                base = buf.ctypes.data
                base += (-base) % 16
                ptr = base + (org[0] * s[0] + org[1] * s[1]) * item
                fa.append((ptr, s, 0, nk, self._vec(ptr, s, item)))
                e = f["extent"] or [[0, 0], [0, 0]]
                pitch, nj = (s[1] if f["dims"][1] else nelem), (dom[1] + e[1][1] - e[1][0] if f["dims"][1] else 1)
                geo[fidx] = (base, [pitch if f["dims"][0] else 1, nj, nk], s, org)
                continue
            arr = fields.get(f["name"])
            if arr is None:
                fa.append((0, [0] * 7, 0, 0, 0))
                continue
            es = [st // arr.itemsize for st in arr.strides]
            org = origins[f["name"]]
            s, o3, shape3, ax = [0] * 7, [0] * 3, [1] * 3, 0
            for a in range(3):
                if f["dims"][a]:
                    s[a], o3[a], shape3[a] = es[ax], int(org[ax]), arr.shape[ax]
                    ax += 1
            for d in range(len(f["data_dims"])):
                s[3 + d] = es[ax + d]
            ptr = arr.ctypes.data + sum(o3[a] * s[a] for a in range(3)) * item
            klo, khi = (-o3[2], shape3[2] - o3[2]) if f["dims"][2] else (0, 1)
            fa.append((ptr, s, klo, khi, self._vec(ptr, s, item)))
            geo[fidx] = (arr.ctypes.data, shape3, s, o3)
        scal = b""
        if plan["scalars"]:
            fmt_of = {"bool": "?", "int8": "b", "int16": "h", "int32": "i", "int64": "q", "float32": "f", "float64": "d"}
            fmt, pos = "<", 0
            vals = []
            for sc in plan["scalars"]:
                if sc["offset"] > pos:
                    fmt += f"{sc['offset'] - pos}x"
                fmt += fmt_of[sc["dtype"]]
                pos = sc["offset"] + b2ir.ITEMSIZE[sc["dtype"]]
                v = params.get(sc["name"], 0)
                vals.append(bool(v) if sc["dtype"] == "bool" else (int(v) if sc["dtype"].startswith("int") else float(v)))
            if plan["scalars_size"] > pos:
                fmt += f"{plan['scalars_size'] - pos}x"
            scal = struct.pack(fmt, *vals)
        i_lo, i_hi, j_lo, j_hi = subbox if subbox is not None else (0, dom[0], 0, dom[1])

        def blob(k_lo, k_hi):
            b = struct.pack("<10i3Q", dom[0], dom[1], dom[2], i_lo, i_hi, j_lo, j_hi, k_lo, k_hi, 0, *halo_wait)  # Geom (64 bytes)
            for ptr, s, klo, khi, vec in (fa if fa else [(0, [0] * 7, 0, 0, 0)]):
                b += struct.pack("<Q7q4i", ptr, *s, klo, khi, vec, 0)
            b += scal
            tmaps = plan.get("tmaps", [])
            if tmaps:  # mirror of launcher.cu encode_tmaps (emulated map layout: b200_device.cuh, B200_HOST_EMU)
                b += b"\0" * ((-len(b)) % 64)
                offs = b""
                for t in tmaps:
                    f = plan["fields"][t["field"]]
                    g, (_p, _s, _klo, _khi, vec) = geo.get(t["field"]), fa[t["field"]]
                    if g is None or not vec:
                        b += b"\0" * 128
                        offs += struct.pack("<4i", 0, 0, 0, 0)
                        continue
                    start, shape3, s, o3 = g
                    item = f["itemsize"]
                    base = start - start % 16
                    extra = (start - base) // item
                    has_k = bool(f["dims"][2])
                    dims = [shape3[0] + extra, shape3[1], shape3[2] if has_k else 1]
                    sk = s[2] * item if has_k else s[1] * item * shape3[1]
                    m = struct.pack("<Q3q3q3i", base, *dims, item, s[1] * item, sk, t["box"][0], t["box"][1], item)
                    b += m + b"\0" * (128 - len(m))
                    offs += struct.pack("<4i", o3[0] + extra, o3[1], o3[2] if has_k else 0, 1 if has_k else 0)
                b += offs
            return b + b"\0" * ((-len(b)) % 8)

        def resolve(bound):
            return bound[1] if bound[0] == "start" else dom[2] + bound[1]

        def launch(k, k_lo, k_hi):
            e = k["extent"]
            nx = (i_hi + e[0][1]) - (i_lo + e[0][0])
            ny = (j_hi + e[1][1]) - (j_lo + e[1][0])
            nz = 1 if k["kind"] == "seq" else k_hi - k_lo
            if nx <= 0 or ny <= 0 or nz <= 0:
                return
            tile = k.get("tile", [k["block"][0], k["block"][1], 1])
            grid = [-(-nx // tile[0]), -(-ny // tile[1]), -(-nz // (tile[2] if k["kind"] != "stream" else 1))]
            if k["kind"] == "stream":
                V = tile[2]
                x0, x1 = i_lo + e[0][0], i_hi + e[0][1]
                qx0 = x0 // V  # floor
                nseg = -(-(x1 - (qx0 - int(k.get("qshift", 0))) * V) // tile[0])
                ntj = -(-ny // tile[1])
                grid = [-(-(nseg * ntj * nz) // k["block"][1]), 1, 1]
            data = blob(k_lo, k_hi)
            raw = ctypes.create_string_buffer(len(data) + 64)  # the argument block is 64-byte aligned (tensor maps)
            addr = ctypes.addressof(raw)
            buf = ctypes.c_void_p(addr + (-addr) % 64)
            ctypes.memmove(buf, data, len(data))
            fn = getattr(self.lib, f"emu_call_{k['name']}")
            g = (ctypes.c_uint * 3)(*grid)
            bdim = (ctypes.c_uint * 3)(*k["block"])
            self.lib.emu_launch(ctypes.cast(fn, ctypes.c_void_p), buf, g, bdim, 1 if k["kind"] == "stream" else 0)
            self.launches += 1

        for step in plan["steps"]:
            if step["t"] == "launch":
                k = plan["kernels"][step["kernel"]]
                launch(k, resolve(k["k_lo"]), resolve(k["k_hi"]))
            else:
                for sec in step["sections"]:
                    k0, k1 = resolve(sec["interval"][0]), resolve(sec["interval"][1])
                    levels = range(k0, k1) if step["order"] == "forward" else range(k1 - 1, k0 - 1, -1)
                    for level in levels:
                        for ki in sec["kernels"]:
                            launch(plan["kernels"][ki], level, level + 1)
        del keep

    @staticmethod
    def _vec(ptr, s, item):
        return int(s[0] == 1 and ptr % 16 == 0 and (s[1] * item) % 16 == 0 and (s[2] * item) % 16 == 0)


# ---- staging in the backend's storage layout, optionally between guard pages ------------------------
_PAGE = mmap.PAGESIZE
_libc = ctypes.CDLL(None, use_errno=True)
_libc.mprotect.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]


def _stage_b200(arr: np.ndarray, origin, guard, keep) -> np.ndarray:
    from gt4py_b200 import storage as b2storage

    dims = b2storage.default_dimensions(arr.ndim)
    lmap = b2storage.layout_map(dims)
    org = tuple(int(o) for o in origin) + (0,) * (arr.ndim - len(origin))
    estrides, total, lead = b2storage.compute_layout(arr.shape, lmap, arr.itemsize, b2storage.ALIGNMENT_ELEMENTS, org)
    nbytes = (total + lead) * arr.itemsize
    span = (nbytes + _PAGE - 1) // _PAGE * _PAGE
    mm = mmap.mmap(-1, span + 2 * _PAGE)
    base = ctypes.addressof(ctypes.c_char.from_buffer(mm))
    for off in (0, _PAGE + span):
        if _libc.mprotect(base + off, _PAGE, 0) != 0:  # PROT_NONE
            raise OSError(ctypes.get_errno(), "mprotect")
    # the allocation handed to the launcher is 256-byte aligned (torch caching allocator): keep that
    # alignment so the vector-path decision is the device's; flush against the end guard if asked
    start = _PAGE if guard != "end" else _PAGE + (span - nbytes) // 256 * 256
    flat = np.frombuffer(mm, dtype=arr.dtype, count=total + lead, offset=start)
    view = np.lib.stride_tricks.as_strided(flat[lead:], shape=arr.shape, strides=tuple(s * arr.itemsize for s in estrides))
    view[...] = arr
    keep.append((mm, flat))
    return view
