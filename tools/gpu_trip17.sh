#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a $O/summary.txt; timeout 900 "$@" > $O/$name.log 2>&1; echo "rc=$? $(tail -n 1 $O/$name.log)" | tee -a $O/summary.txt; }
rm -f $O/summary.txt
run tests       python -m pytest -q -m gpu -p no:cacheprovider tests
DFINE_GEMM=tc run conv_base   python tools/bench_conv.py
run bench_tc3   python bench.py --steps 10 --warmup 3 --no-cpu-baseline
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
run ncu_attn $NCU -k regex:"attn_mma" -c 3 -o $O/prof_r1_attn python tools/profile_step.py --eager
cat $O/summary.txt
