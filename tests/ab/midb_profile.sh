#!/bin/bash
# launch-level timing (ncu gpu__time_duration) of our kernels for mid-size batches
mkdir -p gpurun_out
for B in 512 1024 2048; do
  python tests/bench_profile_case.py $B 5 > gpurun_out/midb_plain_$B.log 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'select|seed|finalize|override|merge' \
      --csv --log-file gpurun_out/midb_launches_$B.csv python tests/bench_profile_case.py $B 1 > gpurun_out/midb_ncu_$B.log 2>&1
done
