#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest -q -m gpu -p no:cacheprovider tests/test_model_gpu.py -k "pretrained" > gpurun_out/tests_m.log 2>&1; echo "rc=$? $(tail -n 1 gpurun_out/tests_m.log)"
grep -E "AssertionError" gpurun_out/tests_m.log | head -5
