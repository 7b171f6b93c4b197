#!/bin/bash
# round-end evidence: bench line, launch list and one --set full capture of the dominant kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/r01_clocks_bench.csv &
SMI=$!
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
kill $SMI
tail -1 gpurun_out/bench_final.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench_final.err
python tests/bench_profile_case.py 4096 1 > gpurun_out/r01_plain_final.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'select|seed|finalize|override|merge' \
    --csv --log-file gpurun_out/r01_launches_final.csv python tests/bench_profile_case.py 4096 1 > gpurun_out/r01_ncu_list_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:select_tc_kernel -s 3 -c 1 \
    -o gpurun_out/r01_select_tc_b4096_hist python tests/bench_profile_case.py 4096 1 > gpurun_out/r01_ncu_full_final.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
