#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest -q -m gpu -p no:cacheprovider tests > gpurun_out/tests.log 2>&1; echo "rc=$? $(tail -n 1 gpurun_out/tests.log)"
grep -E "^FAILED|^E   " gpurun_out/tests.log | head -10
python __graft_entry__.py smoke 2>&1 | tail -1
