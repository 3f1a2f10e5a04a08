#!/bin/sh
# Memory-safety pass without compute-sanitizer (closed on the GPU pool): build libfcpp with device asserts at the
# shared-memory indexers of the hot kernels (-DFCPP_BOUNDS_DEBUG) and run the whole GPU test suite through it.
# A failed assert traps the kernel ("device-side assert triggered") and fails the test that launched it.
#   sh tools/bounds_check.sh            (on a GPU box; ~2 minutes)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
make -s -C "$ROOT/field_coverage_path_planning_b200/csrc" -j4 VARIANT=bounds DEFS="-DFCPP_BOUNDS_DEBUG" > /dev/null
cd "$ROOT"
FCPP_LIB="$ROOT/field_coverage_path_planning_b200/variants/libfcpp_bounds.so" python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^frame" | tail -6
FCPP_LIB="$ROOT/field_coverage_path_planning_b200/variants/libfcpp_bounds.so" python tools/sanitize_case.py
