#!/bin/bash
# Run the REFERENCE's own cartesian test-suites against the b200 code generator on the CPU emulator
# (test-only backend "b200emu", tests/emu/emu_backend.py).  Build container only (needs /root/reference).
#   tools/run_reference_tests.sh            -> test_code_generation.py + test_suites.py + feature_tests/
#                                              + test_math_functions.py + backend / builder unit tests (≈6 min; 198 tests on b200emu)
#   tools/run_reference_tests.sh --real-backend  -> the same suites on the real backend="b200" classes on the fake device
set -e
REPO="$(cd "$(dirname "$0")/.." && pwd)"
WORK="${TMPDIR:-/tmp}/gt4py_b200_reftests"; mkdir -p "$WORK/cache"; cd "$WORK"
export PYTHONPATH="$REPO/tests:$REPO/tools/shims:/root/reference/src:$REPO:/root/reference/tests"
export GT_CACHE_ROOT="$WORK/cache"
I=/root/reference/tests/cartesian_tests/integration_tests
T=$I/multi_feature_tests
U=/root/reference/tests/cartesian_tests/unit_tests
if [ "${1:-}" = "--real-backend" ]; then
  # the REAL plug-in classes (backend="b200": B200Backend, B200StencilObject, storage hooks, DeviceArray) on the fake
  # device (tests/emu/fake_device.py); deselected: the reference tests that need cupy itself (get_array_library)
  WORK="${WORK}_real"; mkdir -p "$WORK/cache"; cd "$WORK"; export GT_CACHE_ROOT="$WORK/cache"
  exec python -m pytest -p emu.fake_device_plugin -p no:cacheprovider --rootdir="$WORK" -c /dev/null -q -W ignore --require-optional-deps \
      "$T/test_suites.py" "$T/test_code_generation.py" "$T/test_math_functions.py" "$I/feature_tests" \
      -k "b200 and not (K_offset_write_simple or K_offset_write_forward or K_offset_write_backward or K_offset_write_conditional or numpy_allocators or bad_layout_warns or data_dimensions_stride_is_always_higher_than_cartesian)"
fi
if [ $# -eq 0 ]; then set -- "$T/test_suites.py" "$T/test_math_functions.py" "$I/feature_tests" \
    "$U/backend_tests/test_backend.py" "$U/backend_tests/test_module_generator.py" "$U/test_stencil_builder.py" "$U/test_lazy_stencil.py"; fi
python -m pytest -p emu.emu_backend_plugin -p no:cacheprovider --rootdir="$WORK" -c /dev/null -q -W ignore \
    "$T/test_code_generation.py" "$@" -k b200emu
