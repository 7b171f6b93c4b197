#!/bin/bash
# round 2, GPU call 15: large batches (pairs + poller vs single CTAs), odd tile counts
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python tests/perf_sweep.py --batches 6144,8192,12288,16384 \
  --variants "default=;pairpo=CCR_2CTA=1,CCR_THR_MODE=1,CCR_THROTTLE=1;pairin=CCR_2CTA=1,CCR_THR_MODE=0,CCR_THROTTLE=1;singlepo=CCR_2CTA=0,CCR_THROTTLE=1,CCR_THR_MODE=1;single=CCR_2CTA=0" \
  --secs 0.5 --rounds 2 --md $O/r02_c15_sweep_large.md > $O/r02_c15_sweep_large.log 2>&1; tail -24 $O/r02_c15_sweep_large.log
timeout 900 python tests/perf_sweep.py --batches 384,640,1152,1664 \
  --variants "default=;thr=CCR_THROTTLE=1;pair=CCR_2CTA=1;pairnothr=CCR_2CTA=1,CCR_THROTTLE=0" \
  --secs 0.4 --rounds 2 --md $O/r02_c15_sweep_odd.md > $O/r02_c15_sweep_odd.log 2>&1; tail -20 $O/r02_c15_sweep_odd.log
