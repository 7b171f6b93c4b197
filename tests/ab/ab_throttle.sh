#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/ab_throttle.txt
for B in 2048 4096 8192; do
  for cfg in "CCR_2CTA=0" "CCR_2CTA=0 CCR_THROTTLE=1" "CCR_2CTA=1" "CCR_2CTA=1 CCR_THROTTLE=1" "CCR_2CTA=0" "CCR_2CTA=1 CCR_THROTTLE=1"; do
    r=$(env $cfg python tests/bench_profile_case.py $B 12 2>&1 | tail -1)
    echo "B=$B [$cfg] $r" | tee -a gpurun_out/ab_throttle.txt
  done
done
