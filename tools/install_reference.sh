#!/usr/bin/env bash
# Populate the git-ignored baseline/_ref/ with the UNMODIFIED reference package (gt4py, pure Python) so that
# it travels to the GPU box with the gpurun snapshot (baseline/_ref is git-ignored, NOT gpurun-ignored):
#
#   * tests -m gpu build stencils with @gtscript.stencil(backend="b200") and call StencilObject.__call__ there,
#   * bench.py --impl reference times the reference's own `numpy` backend there.
#
# Build container only (needs /root/reference).  Never copies anything into tracked paths.
#   1. the contract's offline pip install (fails here: the build back-end needs `versioningit`, which is not in
#      /opt/wheelhouse) -> 2. fall back to copying src/gt4py as it lies (the package is pure Python; its
#      compiled GridTools back-ends are out of scope and not buildable offline).
# The reference's absent pure-Python dependencies are stood in for by the tracked dev shims in tools/shims/
# (boltons, deepdiff, toolz, mako, devtools, cached_property, array_api_compat, black, gridtools_cpp: only the
# handful of functions gt4py.cartesian imports), which tools/refenv.py puts on sys.path next to baseline/_ref.
set -u
REPO="$(cd "$(dirname "$0")/.." && pwd)"
REF="${1:-/root/reference}"
DEST="$REPO/baseline/_ref"
[ -d "$REF/src/gt4py" ] || { echo "install_reference: $REF/src/gt4py not found (GPU box?): nothing to do"; exit 0; }
rm -rf "$DEST"; mkdir -p "$DEST"
TMP="$(mktemp -d)"; trap 'rm -rf "$TMP"' EXIT
cp -r "$REF" "$TMP/src" 2>/dev/null
if python -m pip install -q --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$DEST" "$TMP/src" >"$TMP/pip.log" 2>&1; then
  echo "install_reference: pip install ok -> $DEST"; echo "pip" > "$DEST/.how"
else
  echo "install_reference: pip install failed ($(grep -m1 -o "No module named '[a-z_]*'" "$TMP/pip.log" || echo see log)); copying src/gt4py as it lies"
  rm -rf "$DEST"; mkdir -p "$DEST"
  cp -r "$REF/src/gt4py" "$DEST/gt4py"
  find "$DEST" -name __pycache__ -type d -prune -exec rm -rf {} +
  echo "copy of $REF/src/gt4py ($(cd "$REF" && git rev-parse --short HEAD 2>/dev/null || echo unknown))" > "$DEST/.how"
fi
PYTHONPATH="$REPO/tools" python - <<'PY'
import refenv
assert refenv.enable_gt4py(prefer_installed=True), "gt4py not importable from baseline/_ref"
import gt4py, gt4py.cartesian.gtscript  # noqa
print("install_reference: import ok:", gt4py.__file__)
PY
