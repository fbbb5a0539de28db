"""pytest plugin: `-p emu.fake_device_plugin` puts the fake device (tests/emu/fake_device.py) under a whole test
session BEFORE test modules are collected, and registers `backend="b200"`.  Used to run the REFERENCE's own
suites against the real plug-in classes (B200Backend, B200StencilObject, the storage hooks) where there is no GPU:
    tools/run_reference_tests.sh --real-backend
Test infrastructure only."""
import pytest

from emu import fake_device

_mp = pytest.MonkeyPatch()
fake_device.install(_mp)

import gt4py_b200  # noqa: E402,F401  (registers the backend and the storage hooks)


def pytest_unconfigure(config):
    _mp.undo()
