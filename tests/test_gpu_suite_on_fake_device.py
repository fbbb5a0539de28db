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


def test_autotune_of_stencils_that_update_fields_in_place_on_the_fake_device():
    """B200Stencil.autotune on column solvers (sup / rhs of the Thomas solver are READ_WRITE): every candidate starts from
    a saved copy of the in-place fields, is validated after one application, and the caller's inputs are restored."""
    import numpy as np
    import torch

    if _have_gpu() or os.environ.get("B200_EMULATE_DEVICE") == "1":
        pytest.skip("real device present / already emulating")
    from emu import fake_device as fd

    mp = pytest.MonkeyPatch()
    fd.install(mp)
    try:
        from gt4py_b200 import storage, testing
        from gt4py_b200.stencil import B200Stencil
        from oracle import numpy_oracle

        for name in ("tridiagonal_f64", "fw_wsolve_f32"):
            st = testing.load_ir(name, "default")
            fields, params, origins, domain = testing.make_case_data(st, name, domain=(37, 5, 9), seed=3)
            ref = {k: v.copy() for k, v in fields.items()}
            numpy_oracle.run(st, ref, params, domain, origins)
            dev = {k: storage.from_array(v, aligned_index=origins[k]) for k, v in fields.items()}
            stencil = B200Stencil(st, {"device_sync": False})
            cands = [{}, {"seq_rotate": False}, {"seq_prefetch": 2}, {"fuse_columns": True}, {"seq_cache": False}]
            tuned = stencil.autotune(dev, params, domain=domain, origin=origins, candidates=cands, iters=2, refine=1)
            assert len(tuned) >= 3 and stencil.tune_rejected == []
            for n, v in fields.items():  # the in-place inputs are back
                if st["field_info"][n]["access"] == "READ_WRITE":
                    np.testing.assert_array_equal(dev[n].get(), v, err_msg=f"{name}:{n} not restored")
            stencil(**dev, **params, origin=origins, domain=domain)
            torch.cuda.synchronize()
            for n in testing.written_fields(st):
                box = tuple(slice(o, o + d) for o, d in zip(origins[n], domain))
                np.testing.assert_array_equal(dev[n].get()[box], ref[n][box], err_msg=f"{name}:{n}")
    finally:
        mp.undo()
