#!/bin/bash
# 3xBF16 forward mode: tap-padded planes; parity tests, bench in both modes, launch list in bf3 mode
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a $O/summary.txt; timeout 900 "$@" > $O/$name.log 2>&1; echo "rc=$? $(tail -n 1 $O/$name.log)" | tee -a $O/summary.txt; }
rm -f $O/summary.txt
run tests       python -m pytest -q -m gpu -p no:cacheprovider tests
DFINE_GEMM=bf3 run bench_bf3   python bench.py --steps 10 --warmup 5 --no-cpu-baseline
run bench_tc3   python bench.py --steps 10 --warmup 5 --no-cpu-baseline
DFINE_GEMM=bf3 run ncu_list_bf3 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_bf3.csv python tools/profile_step.py --eager
cat $O/summary.txt
