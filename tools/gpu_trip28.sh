#!/bin/bash
# taps off, ELAN concat restored, dw 3x3 s2 dgrad kernel, shared d(memory) buffer: parity + bench + launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a $O/summary.txt; timeout 900 "$@" > $O/$name.log 2>&1; echo "rc=$? $(tail -n 1 $O/$name.log)" | tee -a $O/summary.txt; }
rm -f $O/summary.txt
run tests       python -m pytest -q -m gpu -p no:cacheprovider tests
run bench_tc3   python bench.py --steps 10 --warmup 5 --no-cpu-baseline
run ncu_list ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches.csv python tools/profile_step.py --eager
cat $O/summary.txt
