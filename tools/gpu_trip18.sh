#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a $O/summary.txt; timeout 900 "$@" > $O/$name.log 2>&1; echo "rc=$? $(tail -n 1 $O/$name.log)" | tee -a $O/summary.txt; }
rm -f $O/summary.txt
run tests       python -m pytest -q -m gpu -p no:cacheprovider tests/test_ops_gpu.py tests/test_model_gpu.py
DFINE_GEMM=tc run conv_box5   python tools/bench_conv.py
DFINE_GEMM=tc DFINE_WGRAD_BOX5=0 run conv_box4   python tools/bench_conv.py
run bench_tc3   python bench.py --steps 10 --warmup 3 --no-cpu-baseline
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
run ncu_attn $NCU -k regex:"^(fwd|dq|dkv)_kernel" -c 3 -o $O/prof_r1_attn python tools/profile_step.py --eager
cat $O/summary.txt
