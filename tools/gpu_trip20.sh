#!/bin/bash
# 3xBF16 forward mode bring-up: accuracy vs fp64, parity tests, bench in both modes; attention issue-order A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a $O/summary.txt; timeout 900 "$@" > $O/$name.log 2>&1; echo "rc=$? $(tail -n 1 $O/$name.log)" | tee -a $O/summary.txt; }
rm -f $O/summary.txt
run diag_bf3    python tools/diag_tf32.py
run tests_ops   python -m pytest -q -m gpu -p no:cacheprovider tests/test_ops_gpu.py -x -k "tc_matches_simt or attention"
run tests_model python -m pytest -q -m gpu -p no:cacheprovider tests/test_model_gpu.py
DFINE_GEMM=bf3 run bench_bf3   python bench.py --steps 10 --warmup 5 --no-cpu-baseline
run bench_tc3   python bench.py --steps 10 --warmup 5 --no-cpu-baseline
DFINE_GEMM=bf3 run conv_bf3 python tools/bench_conv.py
cat $O/summary.txt; tail -5 $O/diag_bf3.log
