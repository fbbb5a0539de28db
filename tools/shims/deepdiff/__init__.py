"""Dev-only stand-in for `deepdiff` (absent offline)."""
from . import deephash, diff
