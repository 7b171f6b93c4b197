#!/bin/bash
# round 2, GPU call 14: same-box A/B inline throttle vs poller-warp throttle
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python tests/perf_sweep.py --batches 512,1024,2048,4096,8192 \
  --variants "in16=CCR_THR_MODE=0,CCR_LEAD=16;in8=CCR_THR_MODE=0,CCR_LEAD=8;po16=CCR_THR_MODE=1,CCR_LEAD=16;po32=CCR_THR_MODE=1,CCR_LEAD=32;po8=CCR_THR_MODE=1,CCR_LEAD=8;nothr=CCR_THROTTLE=0" \
  --secs 0.4 --rounds 3 --md $O/r02_c14_sweep_thr.md > $O/r02_c14_sweep_thr.log 2>&1; tail -40 $O/r02_c14_sweep_thr.log
