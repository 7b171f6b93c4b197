#!/bin/bash
# round 2, GPU call 13: poller-warp throttle -- correctness, lead sweep, DRAM bytes
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_fuzz.py tests/test_gpu_parity.py -m gpu -q --maxfail=10 > $O/r02_c13_pytest.log 2>&1; echo "pytest rc $?" >> $O/r02_c13_pytest.log
tail -5 $O/r02_c13_pytest.log
timeout 900 python tests/perf_sweep.py --batches 512,768,1024,1536,2048,4096,8192 \
  --variants "default=;lead3=CCR_LEAD=3;lead6=CCR_LEAD=6;lead10=CCR_LEAD=10;lead16=CCR_LEAD=16;nothr=CCR_THROTTLE=0" \
  --secs 0.4 --rounds 2 --md $O/r02_c13_sweep_poller.md > $O/r02_c13_sweep_poller.log 2>&1; tail -50 $O/r02_c13_sweep_poller.log
timeout 300 python tests/batch_case.py 512,1024,2048,4096 100 0 > $O/r02_c13_batch_case_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.avg.per_second \
  --clock-control none -k regex:select_tc_kernel --csv --log-file $O/r02_c13_batch_case_ncu.csv \
  python tests/batch_case.py 512,1024,2048,4096 100 0 > $O/r02_c13_batch_case_ncu.log 2>&1
grep "gpu__time_duration\|dram__bytes_read" $O/r02_c13_batch_case_ncu.csv | awk -F'","' '{print $5, $9, $13, $NF}' | grep -v "<3" | cut -c1-160
