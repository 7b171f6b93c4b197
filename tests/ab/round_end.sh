#!/bin/bash
# what the driver runs at round end, plus the ncu launch list of bench.py itself
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
python bench.py --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_round_end.json; cut -c1-330 gpurun_out/bench_round_end.json
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r01_bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'select|seed|finalize|override|merge|ingest' \
    --csv --log-file gpurun_out/r01_launches_bench_py.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r01_ncu_bench_py.log 2>&1
tail -1 gpurun_out/r01_ncu_bench_py.log | cut -c1-200
