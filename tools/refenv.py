"""Make the reference package (gt4py) importable for TESTS and the bench's reference arm — never for the product.

Search order: an already importable gt4py; `baseline/_ref/` (populated by tools/install_reference.sh, travels to
the GPU box); `/root/reference/src` (build container only).  The reference's absent pure-Python dependencies come
from the dev shims in tools/shims/."""

from __future__ import annotations

import importlib.util
import os
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
INSTALLED = ROOT / "baseline" / "_ref"
SOURCE = pathlib.Path("/root/reference/src")


def enable_gt4py(prefer_installed: bool = True) -> bool:
    if os.environ.get("B200_NO_REFERENCE") == "1":  # simulate a box without the reference
        return False
    if importlib.util.find_spec("gt4py") is not None:
        return True
    shims = ROOT / "tools" / "shims"
    cands = [INSTALLED, SOURCE] if prefer_installed else [SOURCE, INSTALLED]
    for c in cands:
        if (c / "gt4py" / "__init__.py").exists() and shims.exists():
            sys.path[:0] = [str(shims), str(c)]
            os.environ.setdefault("GT_CACHE_ROOT", os.path.join(os.environ.get("TMPDIR", "/tmp"), "gt4py_b200_test_cache"))
            return importlib.util.find_spec("gt4py") is not None
    return False


def where() -> str:
    import gt4py

    return str(pathlib.Path(gt4py.__file__).resolve().parent)
