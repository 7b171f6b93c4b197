#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err
tail -1 gpurun_out/bench_final2.json
python tests/bench_profile_case.py 4096 1 > gpurun_out/r01_plain_final2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'select|seed|finalize|override|merge' \
    --csv --log-file gpurun_out/r01_launches_final2.csv python tests/bench_profile_case.py 4096 1 > gpurun_out/r01_ncu_list_final2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:select_tc_kernel -s 3 -c 1 \
    -o gpurun_out/r01_select_tc_b4096_pairs python tests/bench_profile_case.py 4096 1 > gpurun_out/r01_ncu_full_final2.log 2>&1
python tests/batch_sweep.py 2>&1 | tee gpurun_out/r01_batch_sweep_v3.md
python tests/config_cases.py 2>&1 | tee gpurun_out/config_cases_v3.md
