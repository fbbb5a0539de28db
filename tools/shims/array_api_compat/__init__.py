"""Dev-only stand-in for `array_api_compat`."""
import numpy as _np


def array_namespace(*arrays, **kwargs):
    return _np


def is_numpy_array(x):
    return isinstance(x, _np.ndarray)


def is_cupy_array(x):
    return False


def is_array_api_obj(x):
    return isinstance(x, _np.ndarray)
