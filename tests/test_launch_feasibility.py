"""Static launch-feasibility check of every kernel the GPU suite / bench will launch (no GPU needed):
the cubin's resource usage (cuobjdump) against the sm_100a per-CTA limits.  A kernel that compiles but
cannot be launched ("too many resources requested for launch", oversized parameter block) would
otherwise only show up on the device."""

import os
import pathlib
import re
import shutil
import subprocess
import sys

import numpy as np
import pytest

from gt4py_b200 import codegen, jit, testing
from gt4py_b200.stencil import B200Stencil

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or shutil.which("nvcc") is None, reason="needs the CUDA toolkit")

REGS_PER_SM = 65536
MAX_REGS_PER_THREAD = 255
MAX_THREADS_PER_CTA = 1024
MAX_STATIC_SMEM = 48 * 1024
HOT = ("hdiff", "upwind5", "tridiagonal", "vadv", "fw", "laplacian")  # BASELINE.json configs
MAX_PARAM_BYTES = 4096  # the classic limit: stay below it so the launch never depends on the driver's large-parameter support


def _resources(source, opts, name):
    jit.compile_cubin(source, opts, name=name)
    out = subprocess.run(["cuobjdump", "-res-usage", str(jit.cubin_path(source, opts, name=name))], capture_output=True, text=True, check=True).stdout
    res = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+) CONSTANT\[0\]:(\d+)", out):
        res[m.group(1)] = dict(zip(("reg", "stack", "shared", "local", "const0"), map(int, m.groups()[1:])))
    return res


def _check(st, opts, name):
    source, plan = codegen.generate(st, dict(opts))
    res = _resources(source, opts, name)
    nf = len(plan["fields"])
    param_bytes = 40 + 80 * max(nf, 1) + int(plan["scalars_size"])
    assert param_bytes <= MAX_PARAM_BYTES, f"{name}: {param_bytes} bytes of kernel parameters"
    for k in plan["kernels"]:
        r = res[k["name"]]
        threads = int(np.prod(k["block"]))
        assert threads <= MAX_THREADS_PER_CTA, (name, k["name"], threads)
        assert r["reg"] <= MAX_REGS_PER_THREAD
        # registers are allocated per warp in units of 8 per thread
        assert -(-r["reg"] // 8) * 8 * threads <= REGS_PER_SM, f"{name}/{k['name']}: {r['reg']} registers x {threads} threads cannot launch"
        assert r["shared"] + int(k.get("smem", 0)) <= MAX_STATIC_SMEM or k.get("smem", 0) > 0
        # ptxas parks one or two loop-invariant values on the stack outside the steady loops of some streaming
        # kernels (8-24 bytes: STL before, LDL after the loop); anything beyond that in a default
        # benchmark kernel is a real spill.  Register-capped autotune candidates (min_blocks) may spill more.
        capped = "min_blocks" in opts
        if k["kind"] in ("stream", "col") and name.split("_")[0] in HOT and not capped:
            assert r["stack"] <= 32 and r["local"] == 0, f"{name}/{k['name']}: spills ({r})"
        assert r["stack"] <= 256, f"{name}/{k['name']}: heavy spilling ({r})"
    return res


@pytest.mark.parametrize("name", testing.list_cases())
def test_every_fixture_kernel_fits_a_cta(name):
    for variant in ("default", "staged"):
        st = testing.load_ir(name, variant)
        for strategy in ("point", "auto"):
            _check(st, {"strategy": strategy}, codegen._cname(f"{name}.{variant}"))
        if any(loop["order"] != "parallel" for loop in st["loops"]):
            for opts in ({"seq_prefetch": False}, {"seq_cache": False}, {"fuse_columns": True, "seq_prefetch": 2}):
                _check(st, opts, codegen._cname(f"{name}.{variant}"))


@pytest.mark.parametrize("name,pitch", [("hdiff_f32", 1056), ("upwind5_f32", 2080), ("hdiff_f64", 160)])
def test_autotune_candidates_fit_a_cta(name, pitch):
    """every code-generation variant the autotuner may try (stencil.DEFAULT_CANDIDATES)"""
    st = testing.load_ir(name, "staged")
    for cand in B200Stencil.DEFAULT_CANDIDATES:
        cand = dict(cand)
        if cand.get("static_pitch") == "auto":
            cand["static_pitch"] = pitch
        _check(st, {"strategy": "auto", "device_sync": False, **cand}, codegen._cname(name))


def test_fused_kernels_fit_a_cta():
    """cross-stencil fusion doubles the register windows of a kernel: it must still launch and not spill"""
    from gt4py_b200 import fuse

    hd, up = testing.load_ir("hdiff_f32", "staged"), testing.load_ir("upwind5_f32", "staged")
    for st in (fuse.repeat(hd, 2, carry=("in_field", "out_field")), fuse.repeat(hd, 3, carry=("in_field", "out_field")),
               fuse.compose("hdiff_upwind", [(hd, {"out_field": "phi"}), (up, {})], intermediates=["phi"])):  # fmt: skip
        for opts in ({}, {"interior_loop": True, "static_pitch": 1056}):
            res = _check(st, opts, codegen._cname(st["name"]))
            assert all(r["reg"] <= 168 for r in res.values()), res  # >= 3 CTAs of 4 warps per SM


def test_occupancy_cap_option_is_validated_and_reaches_the_plan():
    st = testing.load_ir("tridiagonal_f64", "default")
    _src, plan = codegen.generate(st, {"fuse_columns": True, "seq_prefetch": 4, "seq_smem_pad": 110 * 1024})
    assert [k["smem"] for k in plan["kernels"]] == [110 * 1024] and " 112640 0 0 0\n" in codegen.plan_to_text(plan)  # (smem, the segment-shift flag, shared memory per level / level cap)
    _check(st, {"fuse_columns": True, "seq_prefetch": 4, "seq_smem_pad": 110 * 1024}, "tridiagonal_f64_default")
    with pytest.raises(ValueError, match="seq_smem_pad"):
        codegen.generate(st, {"seq_smem_pad": 300 * 1024})


def test_concurrent_compilation_of_one_variant_never_publishes_a_partial_cubin(tmp_path, monkeypatch):
    """one rank per GPU: the first call of a new variant compiles it in every process at the same time"""
    import concurrent.futures as cf
    import subprocess as sp

    monkeypatch.setenv("GT4PY_B200_CACHE", str(tmp_path))
    src, _plan = codegen.generate(testing.load_ir("copy_f64", "default"), {"strategy": "auto"})
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from gt4py_b200 import jit\n"
        "b = jit.compile_cubin(open(%r).read(), {'strategy': 'auto'}, name='race')\n"
        "print(len(b))\n"
    ) % (str(pathlib.Path(__file__).resolve().parent.parent), str(tmp_path / "src.cu"))
    (tmp_path / "src.cu").write_text(src)
    with cf.ThreadPoolExecutor(6) as ex:
        outs = list(ex.map(lambda _n: sp.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, GT4PY_B200_CACHE=str(tmp_path))), range(6)))
    sizes = {o.stdout.strip() for o in outs}
    assert all(o.returncode == 0 for o in outs), outs[0].stderr[-800:]
    assert len(sizes) == 1 and int(sizes.pop()) > 1000
    cubin = next(tmp_path.glob("race_*.cubin"))
    elf = sp.run(["cuobjdump", "-elf", str(cubin)], capture_output=True, text=True).stdout
    assert "b200_copy_f64" in elf
