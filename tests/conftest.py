import pathlib
import sys

import pytest

# `specialize="lazy"` is the backend's default (a second, per-pitch cubin at the first call): the suites pin "off" so that
# tests which count kernels / compilations see the generic kernels unless they ask for the specialisation themselves
__import__("os").environ.setdefault("GT4PY_B200_SPECIALIZE", "off")

ROOT = pathlib.Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "needs_gt4py: test needs the gt4py frontend (reference) importable")


def _have_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


EMULATE = __import__("os").environ.get("B200_EMULATE_DEVICE") == "1"


@pytest.fixture(scope="session", autouse=True)
def _emulated_device():
    """B200_EMULATE_DEVICE=1: the `-m gpu` tests run on a fake device (tests/emu/fake_device.py) so that their
    own logic can be checked where there is no GPU."""
    if not EMULATE or _have_gpu():
        yield
        return
    from emu import fake_device

    mp = pytest.MonkeyPatch()
    fake_device.install(mp)
    yield
    mp.undo()


def pytest_collection_modifyitems(config, items):
    gpu = _have_gpu()
    if EMULATE and not gpu:
        from emu import fake_device

        skip_emu = pytest.mark.skip(reason="needs real device behaviour (not emulated)")
        for item in items:
            if "gpu" in item.keywords and any(s in item.nodeid for s in fake_device.SKIP):
                item.add_marker(skip_emu)
            elif "gpu" not in item.keywords and any(s in item.nodeid for s in fake_device.SKIP):
                item.add_marker(skip_emu)
        gpu = True
    skip_gpu = pytest.mark.skip(reason="no CUDA device")
    skip_gt = pytest.mark.skip(reason="gt4py frontend not importable here")
    for item in items:
        if "gpu" in item.keywords and not gpu:
            item.add_marker(skip_gpu)
        if "needs_gt4py" in item.keywords and not HAVE_GT4PY:
            item.add_marker(skip_gt)


def _try_enable_gt4py():
    """The gt4py frontend (the reference) is made importable for the plug-in tests from `baseline/_ref/`
    (tools/install_reference.sh; travels to the GPU box) or, in the build container, from /root/reference/src,
    through the dev shims for its missing pure-Python deps (tools/refenv.py)."""
    sys.path.insert(0, str(ROOT / "tools"))
    import refenv

    return refenv.enable_gt4py()


HAVE_GT4PY = _try_enable_gt4py()
