#!/bin/bash
# attention rewrite (2 m-tiles / warp, permuted contraction index) + MSDA forward corner slots: parity, bench, launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a $O/summary.txt; timeout 900 "$@" > $O/$name.log 2>&1; echo "rc=$? $(tail -n 1 $O/$name.log)" | tee -a $O/summary.txt; }
rm -f $O/summary.txt
run tests       python -m pytest -q -m gpu -p no:cacheprovider tests
run bench_tc3   python bench.py --steps 10 --warmup 5 --no-cpu-baseline
run ncu_list ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches.csv python tools/profile_step.py --eager
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
run ncu_attn $NCU -k regex:"^(fwd|dq|dkv)_kernel" -c 3 -s 3 -o $O/prof_r1_attn2 python tools/profile_step.py --eager
run ncu_msda $NCU -k regex:"msda_(fwd|bwd)_kernel" -c 2 -o $O/prof_r1_msda2 python tools/profile_step.py --eager
cat $O/summary.txt
