"""The `-m gpu` tests that no device has run yet, executed on the fake device (tests/emu/fake_device.py):
storages on the host, every launch through the CPU emulator of the generated kernels.  Checks the TESTS
(data flow, argument plumbing, assertions) so that the GPU box does not stop on a bug of a test under `-x`.
The whole GPU suite can be run the same way:  B200_EMULATE_DEVICE=1 python -m pytest tests -m gpu -q"""

import os
import pathlib
import subprocess
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent


def _have_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_have_gpu() or os.environ.get("B200_EMULATE_DEVICE") == "1", reason="real device present / already emulating")
def test_unverified_gpu_tests_pass_on_the_fake_device():
    env = dict(os.environ, B200_EMULATE_DEVICE="1")
    files = ["tests/test_zz_fuse_gpu.py", "tests/test_zz_graph_gpu.py", "tests/test_zzz_capture_error_gpu.py", "tests/test_hostpipe.py",
             "tests/test_storage.py", "tests/test_parity_gpu.py::test_hdiff_smooth_analytic_field"]
    proc = subprocess.run([sys.executable, "-m", "pytest", *files, "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider"],
                          cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)  # fmt: skip
    assert proc.returncode == 0, proc.stdout[-3000:]
    assert " passed" in proc.stdout and "failed" not in proc.stdout
