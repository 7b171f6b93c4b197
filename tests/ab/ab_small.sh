#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/ab_small.txt
for B in 192 256 384 512 1024; do
  for cfg in "CCR_2CTA=0" "CCR_2CTA=1" "CCR_2CTA=0 CCR_NO_HIST=1" "CCR_2CTA=1 CCR_NO_HIST=1"; do
    r=$(env $cfg python tests/bench_profile_case.py $B 10 2>&1 | tail -1)
    echo "B=$B [$cfg] $r" | tee -a gpurun_out/ab_small.txt
  done
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'select|seed|finalize|override|merge' \
      --csv --log-file gpurun_out/b256_launches.csv python tests/bench_profile_case.py 256 1 > gpurun_out/b256_ncu.log 2>&1
