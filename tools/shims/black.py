"""Dev-only stand-in for `black` (absent offline): formatting is the identity."""
import enum


class _TV(enum.Enum):
    PY310 = 10
    PY311 = 11
    PY312 = 12
    PY313 = 13


TargetVersion = _TV


class FileMode:
    def __init__(self, **kwargs):
        self.kwargs = kwargs


Mode = FileMode


def format_str(src, *, mode=None):
    return src
