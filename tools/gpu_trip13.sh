#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a $O/summary.txt; timeout 600 "$@" > $O/$name.log 2>&1; echo "rc=$? $(tail -n 1 $O/$name.log)" | tee -a $O/summary.txt; }
rm -f $O/summary.txt
export DFINE_GEMM=tc
run conv_base   python tools/bench_conv.py
DFINE_TC_DBG=10 run conv_loads_only python tools/bench_conv.py
DFINE_TC_DBG=11 run conv_mma_only python tools/bench_conv.py
DFINE_TC_DBG=12 run conv_w_only python tools/bench_conv.py
DFINE_TC_DBG=13 run conv_a_only python tools/bench_conv.py
DFINE_TC_PERSIST=0 run conv_nonpersist python tools/bench_conv.py
DFINE_GEMM=tc3 run conv_tc3 python tools/bench_conv.py
unset DFINE_GEMM
run tests       python -m pytest -q -m gpu -p no:cacheprovider tests/test_ops_gpu.py
cat $O/summary.txt
