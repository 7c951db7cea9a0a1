#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a $O/summary30.txt; timeout 900 "$@" > $O/$name.log 2>&1; echo "rc=$? $(tail -n 1 $O/$name.log)" | tee -a $O/summary30.txt; }
rm -f $O/summary30.txt
run tests       python -m pytest -q -m gpu -p no:cacheprovider tests
run bench_tc3   python bench.py --steps 20 --warmup 5
bash tools/gpu_trip_final4.sh
cat $O/summary30.txt
