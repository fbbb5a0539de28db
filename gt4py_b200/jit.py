"""JIT driver: CUDA source -> sm_100a cubin with nvcc, cached on disk.

Replaces the reference's setuptools/pybind11/nvcc extension build (reference:
backend/pyext_builder.py:53-167, 324-359; flag conventions utils/compiler.py:154-189): generated
kernels are plain SIMT C++ without template libraries, so a stencil compiles in ~1-2 s to a bare
cubin that the launcher loads through cudaLibraryLoadData.

FMA contraction is OFF by default (``-fmad=false``): the NumPy oracle never fuses, the flux limiter
of horizontal diffusion is discontinuous in the sign of a product (SURVEY §7 hard parts), and the
kernels are HBM-bound so the extra instruction is free.  IEEE division / sqrt, no fast-math.
"""

from __future__ import annotations

import hashlib
import os
import pathlib
import shutil
import subprocess
import tempfile
from typing import Dict, Optional, Sequence

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
CSRC = pathlib.Path(__file__).resolve().parent / "csrc"


def cache_dir() -> pathlib.Path:
    root = os.environ.get("GT4PY_B200_CACHE")
    path = pathlib.Path(root) if root else pathlib.Path(__file__).resolve().parent / "_cache"
    path.mkdir(parents=True, exist_ok=True)
    return path


def nvcc_path() -> str:
    """nvcc location, resolved like the reference resolves its CUDA toolchain (utils/compiler.py:155-189:
    `CUDA_HOME`, then `CUDA_PATH`, then /usr/local/cuda); `NVCC` overrides, `nvcc` on PATH is the last resort."""
    cands = [os.environ.get("NVCC")]
    for var in ("CUDA_HOME", "CUDA_PATH"):
        if os.environ.get(var):
            cands.append(os.path.join(os.environ[var], "bin", "nvcc"))
    cands += ["/usr/local/cuda/bin/nvcc", shutil.which("nvcc")]
    for cand in cands:
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("b200: nvcc not found (set NVCC or CUDA_HOME); no CPU fallback exists")


_NVCC_VERSION: Dict[str, str] = {}


def nvcc_version() -> str:
    """`nvcc --version` text of the resolved compiler (part of every cache key: a toolkit upgrade must not
    serve cubins of the previous compiler)."""
    path = nvcc_path()
    if path not in _NVCC_VERSION:
        try:
            _NVCC_VERSION[path] = subprocess.run([path, "--version"], capture_output=True, text=True).stdout
        except OSError:
            _NVCC_VERSION[path] = "unknown"
    return _NVCC_VERSION[path]


def env_build_settings() -> Dict[str, object]:
    """The reference's build environment (config.py:24-31, utils/compiler.py:171-178), read at call time:
    GT4PY_COMPILE_OPT_LEVEL ("0".."3"), GT4PY_EXTRA_COMPILE_OPT_FLAGS and GT4PY_CARTESIAN_EXTRA_CUDA_COMPILE_ARGS
    (space-separated strings appended to the nvcc command line)."""
    out: Dict[str, object] = {}
    lvl = os.environ.get("GT4PY_COMPILE_OPT_LEVEL")
    if lvl is not None:
        out["opt_level"] = lvl
    extra = " ".join(x for x in (os.environ.get("GT4PY_EXTRA_COMPILE_OPT_FLAGS", ""),
                                 os.environ.get("GT4PY_CARTESIAN_EXTRA_CUDA_COMPILE_ARGS", "")) if x.strip())  # fmt: skip
    if extra:
        out["extra_opt_flags"] = extra
    return out


def compile_flags(options: Optional[Dict] = None) -> list:
    """nvcc flags: explicit options win over the reference's environment variables, which win over the defaults."""
    options = {**env_build_settings(), **{k: v for k, v in (options or {}).items() if v is not None}}
    opt = str(options.get("opt_level", 3))  # the reference passes "0".."3" / "s" as strings (gtc_common.py:192)
    flags = list(ARCH_FLAGS) + ["-std=c++17", "-lineinfo", f"-O{int(opt) if opt.isdigit() else 3}"]
    flags.append("-fmad=true" if options.get("fmad", False) else "-fmad=false")
    if options.get("debug_mode", False):
        flags += ["-G"]
    extra = options.get("extra_opt_flags", []) or []
    flags += extra.split() if isinstance(extra, str) else list(extra)  # reference: one space-separated string
    return flags


def _cubin_key(source: str, flags: Sequence[str]) -> str:
    header = (CSRC / "b200_device.cuh").read_bytes()
    h = hashlib.sha256(source.encode() + b"\0" + " ".join(flags).encode() + b"\0" + header + b"\0" + nvcc_version().encode())
    return h.hexdigest()[:24]


def compile_cubin(source: str, options: Optional[Dict] = None, *, name: str = "stencil", verbose: bool = False) -> bytes:
    """Compile `source` to a cubin (cached by content hash of source + flags + device header)."""
    flags = compile_flags(options)
    key = _cubin_key(source, flags)
    cdir = cache_dir()
    cubin = cdir / f"{name}_{key}.cubin"
    if cubin.exists() and cubin.stat().st_size > 0:
        return cubin.read_bytes()
    cu = cdir / f"{name}_{key}.cu"
    # several processes may compile the same variant at the same time (one rank per GPU, first call of a new variant):
    # the source is published with an atomic rename (same content from every process, so the file nvcc reads is always
    # complete) and every process compiles into its own output — a shared .cu that one rank is still writing would be
    # an empty translation unit for another rank's nvcc, and the kernel-less cubin of that run would land in the cache
    # ("named symbol not found", 8-GPU session r02q)
    with tempfile.NamedTemporaryFile(dir=cdir, suffix=".cu.tmp", delete=False, mode="w", encoding="utf-8") as src_tmp:
        src_tmp.write(source)
    os.replace(src_tmp.name, cu)
    with tempfile.NamedTemporaryFile(dir=cdir, suffix=".cubin", delete=False) as tmp:
        tmp_path = tmp.name
    cmd = [nvcc_path(), *flags, "-cubin", "-I", str(CSRC), str(cu), "-o", tmp_path]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        try:
            os.unlink(tmp_path)
        except OSError:
            pass
        raise RuntimeError(f"b200: nvcc failed for {name}:\n{proc.stderr[-4000:]}\n(source kept at {cu})")
    if verbose and proc.stderr:
        print(proc.stderr)
    os.replace(tmp_path, cubin)
    return cubin.read_bytes()


_GEN_FP = None


def generator_fingerprint() -> str:
    """Hash of everything that decides the generated device code besides the stencil IR and the
    options: the code generators, the device header, the compiler.  Persisted artefacts (cubin +
    launch plan in gt4py's .gt_cache, see backend.py) are only reused under the same fingerprint."""
    global _GEN_FP
    if _GEN_FP is None:
        h = hashlib.sha256()
        pkg = CSRC.parent
        for f in ("codegen.py", "codegen_stream.py", "codegen_column.py", "ir.py", "csrc/b200_device.cuh"):
            h.update((pkg / f).read_bytes())
        try:
            h.update(nvcc_version().encode())
        except Exception:
            h.update(b"no-nvcc")
        _GEN_FP = h.hexdigest()[:20]
    return _GEN_FP


def cubin_path(source: str, options: Optional[Dict] = None, *, name: str = "stencil") -> pathlib.Path:
    return cache_dir() / f"{name}_{_cubin_key(source, compile_flags(options))}.cubin"


def build_launcher(force: bool = False) -> pathlib.Path:
    """Build libgt4py_b200.so in-tree (nvcc -shared, sm_100a)."""
    out = CSRC.parent / "libgt4py_b200.so"
    srcs = [CSRC / "launcher.cu"]
    deps = srcs + [CSRC.parent.parent / "include" / "gt4py_b200.h"]
    if not force and out.exists() and all(out.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return out
    cmd = [
        nvcc_path(), *ARCH_FLAGS, "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
        "-cudart", "static", *[str(s) for s in srcs], "-o", str(out), "-ldl",
    ]  # fmt: skip
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"b200: building the launcher failed:\n{proc.stderr[-4000:]}")
    return out
