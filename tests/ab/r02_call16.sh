#!/bin/bash
# round 2, GPU call 16: final-policy validation -- full GPU suite, batch sweep, bench (both arms), --set full capture
mkdir -p gpurun_out
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=10 > $O/r02_c16_pytest.log 2>&1; echo "pytest rc $?" >> $O/r02_c16_pytest.log
tail -5 $O/r02_c16_pytest.log
timeout 900 python tests/perf_sweep.py --batches 1,8,32,128,256,384,512,768,1024,1536,2048,3072,4096,8192,16384 \
  --variants "default=" --secs 0.5 --rounds 3 --md $O/r02_c16_batch_sweep.md > $O/r02_c16_batch_sweep.log 2>&1; tail -17 $O/r02_c16_batch_sweep.log
timeout 600 python tests/perf_sweep.py --batches 3452 --n 2681468 --k 1001 --variants "default=" --secs 0.5 --rounds 2 --md $O/r02_c16_nq.md > $O/r02_c16_nq.log 2>&1; tail -1 $O/r02_c16_nq.log
timeout 600 python tests/perf_sweep.py --batches 256,4096 --n 12500000 --k 1000 --mask 1 --variants "default=" --secs 0.5 --rounds 2 --md $O/r02_c16_c5shard.md > $O/r02_c16_c5shard.log 2>&1; tail -2 $O/r02_c16_c5shard.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/r02_c16_bench.json 2> $O/r02_c16_bench.err; tail -c 2600 $O/r02_c16_bench.json; tail -2 $O/r02_c16_bench.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > $O/r02_c16_bench_ref.json 2> $O/r02_c16_bench_ref.err; tail -c 600 $O/r02_c16_bench_ref.json; tail -2 $O/r02_c16_bench_ref.err
timeout 300 python tests/batch_case.py 4096 100 1 > $O/r02_c16_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:select_tc_kernel -s 3 -c 1 \
  -o $O/r02_select_tc_b4096_final python tests/batch_case.py 4096 100 1 > $O/r02_c16_ncu_full.log 2>&1
tail -2 $O/r02_c16_ncu_full.log
