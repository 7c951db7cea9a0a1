#!/bin/bash
# hybrid tf32 + bf16-cross-term forward mode (tch): accuracy, parity, bench vs tc3 / bf3 (8 converter warps)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a $O/summary.txt; timeout 900 "$@" > $O/$name.log 2>&1; echo "rc=$? $(tail -n 1 $O/$name.log)" | tee -a $O/summary.txt; }
rm -f $O/summary.txt
run diag        python tools/diag_tf32.py
run tests       python -m pytest -q -m gpu -p no:cacheprovider tests
DFINE_GEMM=tch run bench_tch   python bench.py --steps 10 --warmup 5 --no-cpu-baseline
DFINE_GEMM=bf3 run bench_bf3   python bench.py --steps 10 --warmup 5 --no-cpu-baseline
run bench_tc3   python bench.py --steps 10 --warmup 5 --no-cpu-baseline
cat $O/summary.txt
