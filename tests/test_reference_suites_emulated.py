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
