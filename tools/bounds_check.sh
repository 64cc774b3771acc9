#!/bin/bash
# compute-sanitizer is closed on the pool this repo is developed on.  Its stand-in: a debug build of the
# library whose kernels check every arena / boundary-buffer / string-slot access against its extent
# (-DTAXI_BOUNDS_CHECK, common.cuh TAXI_CHECK) and flag violations in the sticky device status, run
# through one launch of every kernel variant and the GPU parity tests of the aligner.
# Usage (on a GPU box): tools/bounds_check.sh > profiles/bounds_check_rNN.log
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
cd "$ROOT"
[ -f variants/lib_bounds.so ] || make -C taxi2_b200/csrc bounds
export TAXI2_B200_LIB=$ROOT/variants/lib_bounds.so
echo "library: $TAXI2_B200_LIB (built with -DTAXI_BOUNDS_CHECK)"
python tools/sanitize_run.py
python -m pytest tests/test_gpu_align.py tests/test_gpu_multi.py tests/test_gpu_api.py -m gpu -q 2>&1 | tail -4
