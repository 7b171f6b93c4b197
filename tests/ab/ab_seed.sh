#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/ab_seed.txt
for B in 4096 1024; do
  for cfg in "CCR_X=0" "CCR_SEED_M=16384" "CCR_SEED_M=8192" "CCR_SEED_M=4096" "CCR_X=0" "CCR_SEED_M=8192"; do
    r=$(env $cfg python tests/bench_profile_case.py $B 12 2>&1 | tail -1)
    echo "B=$B [$cfg] $r" | tee -a gpurun_out/ab_seed.txt
  done
done
for cfg in "CCR_X=0" "CCR_SEED_M=8192"; do
  echo "## $cfg" | tee -a gpurun_out/ab_seed.txt
  env $cfg python tests/config_cases.py 2>&1 | grep "^| C" | tee -a gpurun_out/ab_seed.txt
done
