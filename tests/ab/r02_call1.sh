#!/bin/bash
# round 2, GPU call 1: new parity tests, reference-CUDA goldens, prefetch sweep, small/mid-batch counters
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 > $O/r02_c1_pytest.log 2>&1; echo "pytest rc $?" >> $O/r02_c1_pytest.log
tail -5 $O/r02_c1_pytest.log
CCR_REFERENCE_ROOT=baseline/_ref/reference timeout 300 python tests/golden/make_golden_cuda.py $O/ > $O/r02_c1_golden.log 2>&1; tail -3 $O/r02_c1_golden.log
timeout 900 python tests/perf_sweep.py --batches 128,256,384,512,1024,2048,4096,8192,16384 \
  --variants "base=;pf2=CCR_PREFETCH=2;pf4=CCR_PREFETCH=4;pf8=CCR_PREFETCH=8;pf16=CCR_PREFETCH=16" \
  --secs 0.3 --rounds 2 --md $O/r02_c1_sweep_prefetch.md > $O/r02_c1_sweep_prefetch.log 2>&1; tail -50 $O/r02_c1_sweep_prefetch.log
timeout 600 python tests/perf_sweep.py --batches 512,4096 --mask 1 \
  --variants "base=;pf4=CCR_PREFETCH=4;pf8=CCR_PREFETCH=8;hq=CCR_HINT_Q=0x14F0000000000000;hi=CCR_HINT_ITEMS=0x12F0000000000000;hqpf=CCR_HINT_Q=0x14F0000000000000,CCR_PREFETCH=4;single=CCR_2CTA=0;singlepf=CCR_2CTA=0,CCR_PREFETCH=4;nothr=CCR_THROTTLE=0,CCR_PREFETCH=4;lead8=CCR_LEAD=8,CCR_PREFETCH=4;lead32=CCR_LEAD=32,CCR_PREFETCH=8" \
  --secs 0.5 --rounds 2 --md $O/r02_c1_sweep_mask.md > $O/r02_c1_sweep_mask.log 2>&1; tail -30 $O/r02_c1_sweep_mask.log
timeout 300 python tests/batch_case.py 1,8,128,256,512,1024 > $O/r02_c1_batch_case_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct \
  --clock-control none -k regex:'select|seed|finalize|override' --csv --log-file $O/r02_c1_batch_case_ncu.csv \
  python tests/batch_case.py 1,8,128,256,512,1024 > $O/r02_c1_batch_case_ncu.log 2>&1
tail -3 $O/r02_c1_batch_case_ncu.log
