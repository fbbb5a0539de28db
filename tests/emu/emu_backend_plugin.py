"""pytest plugin: register the test-only `b200emu` backend before reference test modules are collected."""
import gt4py_b200  # noqa: F401

from emu import emu_backend  # noqa: F401
