#!/bin/bash
# round 2, GPU call 4: full GPU suite after the override fix, default-policy batch sweep, DRAM counters
mkdir -p gpurun_out
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=15 > $O/r02_c4_pytest.log 2>&1; echo "pytest rc $?" >> $O/r02_c4_pytest.log
tail -8 $O/r02_c4_pytest.log
timeout 900 python tests/perf_sweep.py --batches 1,8,32,128,256,384,512,768,1024,1536,2048,3072,4096,8192,16384 \
  --variants "default=" --secs 0.5 --rounds 3 --md $O/r02_c4_batch_sweep.md > $O/r02_c4_batch_sweep.log 2>&1; tail -20 $O/r02_c4_batch_sweep.log
timeout 600 python tests/perf_sweep.py --batches 3452 --n 2681468 --k 1001 --variants "default=" --secs 0.5 --rounds 2 --md $O/r02_c4_nq.md > $O/r02_c4_nq.log 2>&1; tail -2 $O/r02_c4_nq.log
timeout 600 python tests/perf_sweep.py --batches 256,4096 --n 12500000 --k 1000 --mask 1 --variants "default=" --secs 0.5 --rounds 2 --md $O/r02_c4_c5shard.md > $O/r02_c4_c5shard.log 2>&1; tail -3 $O/r02_c4_c5shard.log
timeout 300 python tests/batch_case.py 1024,2048,4096 100 1 > $O/r02_c4_batch_case_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.avg.per_second \
  --clock-control none -k regex:'select|seed|finalize|override' --csv --log-file $O/r02_c4_batch_case_ncu.csv \
  python tests/batch_case.py 1024,2048,4096 100 1 > $O/r02_c4_batch_case_ncu.log 2>&1
tail -3 $O/r02_c4_batch_case_ncu.log
