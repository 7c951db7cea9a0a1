#!/bin/bash
# End-of-round evidence run (session 6): tests, smoke, both bench arms, optional modes, launch list, DRAM traffic of the
# dominant kernel families, --set full captures of the top kernels.  tools/ncu_summary.py turns them into profiles/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a $O/summary.txt; timeout 1200 "$@" > $O/$name.log 2>&1; echo "rc=$? $(tail -n 1 $O/$name.log)" | tee -a $O/summary.txt; }
rm -f $O/summary.txt
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/gpu.txt 2>&1
run tests       python -m pytest -q -m gpu -p no:cacheprovider tests
run smoke       python __graft_entry__.py smoke
run bench_ref   python bench.py --impl reference --steps 2 --warmup 1
run bench_tc3   python bench.py --steps 20 --warmup 5
DFINE_GEMM=tc  run bench_tc  python bench.py --steps 20 --warmup 5 --no-cpu-baseline
DFINE_GEMM=bf3 run bench_bf3 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
DFINE_WGRAD_STREAM=0 run bench_nows python bench.py --steps 20 --warmup 5 --no-cpu-baseline
run ncu_list    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches.csv python tools/profile_step.py --eager
run ncu_traffic ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -k regex:"tc_fwd_persist|tc_wgrad_kernel|msda_" --csv --log-file $O/traffic.csv python tools/profile_step.py --eager
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
run ncu_persist $NCU -k regex:"tc_fwd_persist" -s 60 -c 8 -o $O/prof_s6_persist python tools/profile_step.py --eager
run ncu_wgrad   $NCU -k regex:"tc_wgrad_kernel" -s 20 -c 4 -o $O/prof_s6_wgrad python tools/profile_step.py --eager
run ncu_msda    $NCU -k regex:"msda_" -c 8 -o $O/prof_s6_msda python tools/profile_step.py --eager
run ncu_attn    $NCU -k regex:"^(fwd|dq|dkv)_kernel" -s 3 -c 3 -o $O/prof_s6_attn python tools/profile_step.py --eager
run ncu_small   $NCU -k regex:"fdr_head|stem_|dwconv3x3s2|bn_finalize_apply|matcher" -c 8 -o $O/prof_s6_small python tools/profile_step.py --eager
cat $O/summary.txt
