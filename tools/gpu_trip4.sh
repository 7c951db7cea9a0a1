#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a $O/summary.txt; timeout 900 "$@" > $O/$name.log 2>&1; echo "rc=$? $(tail -n 1 $O/$name.log)" | tee -a $O/summary.txt; }
rm -f $O/summary.txt
PT="python -m pytest -q -m gpu -p no:cacheprovider"
run model       $PT tests/test_model_gpu.py
run bench       python bench.py --steps 10 --warmup 3
cat $O/summary.txt
