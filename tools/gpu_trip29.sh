#!/bin/bash
# 128-byte-aligned pixel strides for the small-channel stem tensors, l / x parity in simt + tc3, x's wide dw layers
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a $O/summary.txt; timeout 900 "$@" > $O/$name.log 2>&1; echo "rc=$? $(tail -n 1 $O/$name.log)" | tee -a $O/summary.txt; }
rm -f $O/summary.txt
run tests       python -m pytest -q -m gpu -p no:cacheprovider tests
run bench_tc3   python bench.py --steps 10 --warmup 5 --no-cpu-baseline
run ncu_list ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches.csv python tools/profile_step.py --eager
python tools/ncu_summary.py launches $O/launches.csv $O/launches.md; rm -f $O/launches.csv
cat $O/summary.txt
