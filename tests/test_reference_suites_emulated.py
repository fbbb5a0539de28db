"""The REFERENCE's own integration test-suite `test_code_generation.py` (every feature test it runs on
every backend: lower-dimensional fields, data dimensions, variable-K offsets, K-offset writes, while
loops, horizontal regions, tables, negative origins, enums, …) executed against the b200 code
generator through the test-only emulated backend `b200emu` (tests/emu/emu_backend.py).
`tools/run_reference_tests.sh` additionally runs the hypothesis suites of test_suites.py (100 tests)."""

import os
import pathlib
import subprocess
import sys
import tempfile

import pytest

pytestmark = pytest.mark.needs_gt4py
ROOT = pathlib.Path(__file__).resolve().parent.parent
REF = pathlib.Path("/root/reference/tests/cartesian_tests/integration_tests/multi_feature_tests")


@pytest.mark.skipif(not REF.exists(), reason="reference tree not mounted")
def test_reference_test_code_generation_passes_on_emulated_b200():
    work = pathlib.Path(tempfile.gettempdir()) / "gt4py_b200_reftests"
    (work / "cache").mkdir(parents=True, exist_ok=True)
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join(
        [str(ROOT / "tests"), str(ROOT / "tools" / "shims"), "/root/reference/src", str(ROOT), "/root/reference/tests"]
    )
    env["GT_CACHE_ROOT"] = str(work / "cache")
    cmd = [
        sys.executable, "-m", "pytest", "-p", "emu.emu_backend_plugin", "-p", "no:cacheprovider", f"--rootdir={work}",
        "-c", "/dev/null", "-q", "-W", "ignore", str(REF / "test_code_generation.py"), "-k", "b200emu",
    ]  # fmt: skip
    proc = subprocess.run(cmd, cwd=work, env=env, capture_output=True, text=True, timeout=1500)
    tail = proc.stdout[-1500:]
    assert proc.returncode == 0, tail
    assert " passed" in tail and "failed" not in tail, tail


# Tests of the reference that need cupy itself (`definitions.get_array_library` asserts it for GPU backends): the one
# divergence of a cupy-free GPU backend; everything else of these files passes on the real plug-in classes.
NEEDS_CUPY = ("K_offset_write_simple or K_offset_write_forward or K_offset_write_backward or K_offset_write_conditional or "
              "numpy_allocators or bad_layout_warns or data_dimensions_stride_is_always_higher_than_cartesian")  # fmt: skip


@pytest.mark.skipif(not REF.exists(), reason="reference tree not mounted")
def test_reference_feature_tests_pass_on_the_real_b200_backend_on_the_fake_device():
    """The reference's feature tests (call interface, exec_info, field layouts, stencil object) parametrised with
    the REAL `backend="b200"` — B200Backend, B200StencilObject, the storage hooks, DeviceArray — running on the fake
    device (tests/emu/fake_device_plugin.py: host memory, launches through the kernel emulator).
    `tools/run_reference_tests.sh --real-backend` runs test_suites.py, test_code_generation.py and
    test_math_functions.py and TestExecInfo the same way (178 tests in all)."""
    work = pathlib.Path(tempfile.gettempdir()) / "gt4py_b200_reftests_real"
    (work / "cache").mkdir(parents=True, exist_ok=True)
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join(
        [str(ROOT / "tests"), str(ROOT / "tools" / "shims"), "/root/reference/src", str(ROOT), "/root/reference/tests"]
    )
    env["GT_CACHE_ROOT"] = str(work / "cache")
    cmd = [
        sys.executable, "-m", "pytest", "-p", "emu.fake_device_plugin", "-p", "no:cacheprovider", f"--rootdir={work}",
        "-c", "/dev/null", "-q", "-W", "ignore", "--require-optional-deps", str(REF.parent / "feature_tests"),
        "-k", f"b200 and not TestExecInfo and not ({NEEDS_CUPY})",  # (TestExecInfo: 75 s on the emulator; in the tools script)
    ]  # fmt: skip
    proc = subprocess.run(cmd, cwd=work, env=env, capture_output=True, text=True, timeout=1500)
    tail = proc.stdout[-1500:]
    assert proc.returncode == 0, tail
    assert " passed" in tail and "failed" not in tail, tail
