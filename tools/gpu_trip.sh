#!/bin/bash
# One parameterised GPU trip (replaces round 1's per-trip scripts):  gpurun -- 'bash tools/gpu_trip.sh <step> [<step> ...]'
# Every step writes gpurun_out/<step>.log; a failing step does not stop the others.  Steps:
#   tests[:<pytest -k expr>]  pytest -m gpu            bench[:<config>]   bench.py (10 steps)
#   smoke                     __graft_entry__.smoke()   ref                bench.py --impl reference (2 steps)
#   tf32peak                  tools/measure_tf32_peak.py
#   refgpu[:<config>]         tools/ref_on_gpu.py
#   conv[:<mode>[:<dbg>]]     tools/bench_conv.py with DFINE_GEMM=<mode> DFINE_TC_DBG=<dbg>
#   traffic[:<config>]        ncu DRAM bytes per kernel family of 1 eager step -> gpurun_out/traffic_<config>.json
#   launches[:<config>]       ncu launch list of 1 eager step -> gpurun_out/launches_<config>.csv + .md summary
#   ncu:<kernel regex>[:<config>]  ncu --set full on up to 3 launches -> summarised csv (the .ncu-rep stays on the box)
#   py:<script.py and args with , for spaces>
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1${TAG:+_$TAG}; shift; echo "== $name: $*" ; ( time timeout -k 20 ${STEP_TIMEOUT:-900} "$@" ) > gpurun_out/$name.log 2>&1; echo "   rc=$? $(tail -n 1 gpurun_out/$name.log | cut -c1-200)"; }
for step in "$@"; do
  IFS=: read -r kind a b c <<< "$step"
  case $kind in
    tests)    if [ -n "$a" ]; then run tests_${b:-sel} python -m pytest tests -q -m gpu -s -k "$a"; else run tests python -m pytest tests -q -m gpu; fi ;;
    smoke)    run smoke python __graft_entry__.py smoke ;;
    bench)    run bench_${a:-m640}${b:+_$b} env ${b:+DFINE_GEMM=$b} python bench.py --config ${a:-m640} --steps 10 --warmup 3 ${c:+--no-cpu-baseline} --dump-launches gpurun_out/shapes_${a:-m640}${TAG:+_$TAG}.md ;;
    ref)      run bench_ref python bench.py --impl reference --steps 2 --warmup 1 ;;
    tf32peak) run tf32peak python tools/measure_tf32_peak.py ;;
    refgpu)   run refgpu_${a:-m640} python tools/ref_on_gpu.py --config ${a:-m640} ;;
    conv)     run conv_${a:-tc3}_${b:-0} env DFINE_GEMM=${a:-tc3} DFINE_TC_DBG=${b:-0} python tools/bench_conv.py ;;
    launches) run launches_${a:-m640} ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_${a:-m640}.csv python tools/profile_step.py --config ${a:-m640}
              python tools/ncu_summary.py launches gpurun_out/launches_${a:-m640}.csv gpurun_out/launches_${a:-m640}.md ;;
    traffic)  run traffic_${a:-m640} ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/traffic_${a:-m640}.csv python tools/profile_step.py --config ${a:-m640}
              python tools/ncu_summary.py traffic gpurun_out/traffic_${a:-m640}.csv gpurun_out/traffic_${a:-m640}.json ;;
    ncu)      tag=$(echo "$a" | tr -c 'A-Za-z0-9' '_'); run ncu_$tag ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$a" -c ${c:-3} -o /tmp/prof_$tag -f python tools/profile_step.py --config ${b:-m640}
              ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$tag.csv 2>/dev/null
              python tools/ncu_summary.py full /tmp/prof_$tag.ncu-rep gpurun_out/ncu_full_$tag.csv ;;
    ncusrc)   # ncusrc:<kernel regex>:<python script and args, "," for spaces>  -> raw + source-page csv of ONE launch
              tag=$(echo "$a$b" | tr -c 'A-Za-z0-9' '_' | cut -c1-120); run ncusrc_$tag ncu --set full --clock-control none --import-source on -k "regex:$a" -s 2 -c 1 -o /tmp/src_$tag -f python ${b//,/ }
              ncu -i /tmp/src_$tag.ncu-rep --page raw --csv > gpurun_out/ncusrc_raw_$tag.csv 2>/dev/null
              ncu -i /tmp/src_$tag.ncu-rep --page source --csv > gpurun_out/ncusrc_source_$tag.csv 2>/dev/null ;;
    py)       tag=$(echo "$a" | tr -c 'A-Za-z0-9' '_' | cut -c1-100); run py_$tag python ${a//,/ } ;;
    *)        echo "unknown step $step" ;;
  esac
done
ls -la gpurun_out | tail -n 40
