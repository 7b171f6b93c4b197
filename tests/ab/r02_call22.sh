#!/bin/bash
# round 2, GPU call 22: validation of the final build -- full GPU suite, bench.py both arms at the driver's
# --steps 20 --warmup 5, ncu launch list of bench.py
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q --maxfail=10 ) > $O/r02_c22_pytest.log 2>&1; echo "pytest rc $?" >> $O/r02_c22_pytest.log
tail -8 $O/r02_c22_pytest.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $O/r02_c22_bench.json 2> $O/r02_c22_bench.err; tail -c 3000 $O/r02_c22_bench.json; tail -4 $O/r02_c22_bench.err
( time timeout 1200 python bench.py --impl reference --steps 20 --warmup 5 ) > $O/r02_c22_bench_ref.json 2> $O/r02_c22_bench_ref.err; tail -c 1500 $O/r02_c22_bench_ref.json; tail -4 $O/r02_c22_bench_ref.err
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-library-baseline > $O/r02_c22_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 -k regex:'ccr|select_|finalize|seed_|override|merge|ingest' --csv \
  --log-file $O/r02_launches_bench_py.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-library-baseline > $O/r02_c22_ncu.log 2>&1
tail -2 $O/r02_c22_ncu.log; wc -l $O/r02_launches_bench_py.csv
