#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 50 python -m pytest -q -m gpu -p no:cacheprovider tests/test_zz_segmentation_gpu.py -rxX > gpurun_out/tests_seg.log 2>&1; echo "rc=$? $(tail -n 1 gpurun_out/tests_seg.log)"
grep -E "XPASS|XFAIL|Error|error" gpurun_out/tests_seg.log | head -12
