#!/bin/bash
# mid-batch A/B: default vs selection-off (CCR_DEBUG=1) vs forced CTA pairs (CCR_2CTA=1)
mkdir -p gpurun_out
: > gpurun_out/midb_ab.txt
for B in 512 1024 2048; do
  for cfg in "" "CCR_DEBUG=1" "CCR_2CTA=1" "CCR_2CTA=1 CCR_DEBUG=1" "CCR_THROTTLE=1"; do
    r=$(env $cfg python tests/bench_profile_case.py $B 8 2>&1 | tail -1)
    echo "B=$B [$cfg] $r" | tee -a gpurun_out/midb_ab.txt
  done
done
