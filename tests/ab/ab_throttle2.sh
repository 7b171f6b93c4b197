#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/ab_throttle2.txt
for B in 512 1024; do
  for cfg in "CCR_2CTA=1 CCR_THROTTLE=0" "CCR_2CTA=1" "CCR_2CTA=0" "CCR_2CTA=1 CCR_THROTTLE=0" "CCR_2CTA=1"; do
    r=$(env $cfg python tests/bench_profile_case.py $B 12 2>&1 | tail -1)
    echo "B=$B [$cfg] $r" | tee -a gpurun_out/ab_throttle2.txt
  done
done
for L in 4 8 16 32 64; do
  r=$(env CCR_2CTA=1 CCR_LEAD=$L python tests/bench_profile_case.py 4096 12 2>&1 | tail -1)
  echo "B=4096 [2CTA lead=$L] $r" | tee -a gpurun_out/ab_throttle2.txt
done
for cfg in "CCR_2CTA=0" "CCR_2CTA=1" "CCR_2CTA=0" "CCR_2CTA=1"; do
  r=$(env $cfg python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['clocks'])")
  echo "bench.py [$cfg] $r" | tee -a gpurun_out/ab_throttle2.txt
done
