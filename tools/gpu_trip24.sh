#!/bin/bash
# weight gradients on a second stream (fork/join inside graph B): parity, graph-vs-eager, bench A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a $O/summary.txt; timeout 900 "$@" > $O/$name.log 2>&1; echo "rc=$? $(tail -n 1 $O/$name.log)" | tee -a $O/summary.txt; }
rm -f $O/summary.txt
run tests       python -m pytest -q -m gpu -p no:cacheprovider tests
run bench_ws    python bench.py --steps 10 --warmup 5 --no-cpu-baseline
DFINE_WGRAD_STREAM=0 run bench_nows  python bench.py --steps 10 --warmup 5 --no-cpu-baseline
cat $O/summary.txt
