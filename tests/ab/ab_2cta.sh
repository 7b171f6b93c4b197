#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/ab_2cta_hist.txt
for B in 2048 3072 4096 8192; do
  for cfg in "CCR_2CTA=0" "CCR_2CTA=1" "CCR_2CTA=0" "CCR_2CTA=1"; do
    r=$(env $cfg python tests/bench_profile_case.py $B 10 2>&1 | tail -1)
    echo "B=$B [$cfg] $r" | tee -a gpurun_out/ab_2cta_hist.txt
  done
done
