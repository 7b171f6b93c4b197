#!/bin/bash
# A/B of the per-row histogram threshold sharing (CCR_NO_HIST=1 disables it), same box
mkdir -p gpurun_out
: > gpurun_out/hist_ab.txt
for cfg in "CCR_NO_HIST=1" "CCR_X=0"; do
  echo "## $cfg" | tee -a gpurun_out/hist_ab.txt
  env $cfg python tests/config_cases.py 2>&1 | grep "^| C" | tee -a gpurun_out/hist_ab.txt
  for B in 512 2048 4096; do
    r=$(env $cfg python tests/bench_profile_case.py $B 10 2>&1 | tail -1); echo "masked bench case B=$B: $r" | tee -a gpurun_out/hist_ab.txt
  done
done
