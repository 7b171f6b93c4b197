#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/midb_ab2.txt
for B in 3072 4096 8192; do
  for cfg in "CCR_2CTA=0" "CCR_2CTA=1" "CCR_2CTA=0" "CCR_2CTA=1"; do
    r=$(env $cfg python tests/bench_profile_case.py $B 10 2>&1 | tail -1)
    echo "B=$B [$cfg] $r" | tee -a gpurun_out/midb_ab2.txt
  done
done
for B in 256 512 1024 2048; do
  r=$(python tests/bench_profile_case.py $B 10 2>&1 | tail -1); echo "B=$B [default] $r" | tee -a gpurun_out/midb_ab2.txt
done
for B in 512 1024; do
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'select|seed|finalize|override|merge' \
      --csv --log-file gpurun_out/midb2_launches_$B.csv python tests/bench_profile_case.py $B 1 > gpurun_out/midb2_ncu_$B.log 2>&1
done
