#!/bin/bash
# round 2, GPU call 7: new ABI entries (argsort, dense tiles, first-hit, BM25 warp kernel), BM25 throughput,
# tiny-table kernel choice, per-kernel times of the k=1000 shard
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bm25.py tests/test_gpu_multi.py -m gpu -q --maxfail=15 > $O/r02_c7_pytest.log 2>&1; echo "pytest rc $?" >> $O/r02_c7_pytest.log
tail -8 $O/r02_c7_pytest.log
timeout 600 python tests/bm25_bench.py > $O/r02_c7_bm25_warp.json 2> $O/r02_c7_bm25_warp.err; tail -c 1200 $O/r02_c7_bm25_warp.json; tail -2 $O/r02_c7_bm25_warp.err
BM25_KERNELS=blockwide timeout 600 python tests/bm25_bench.py > $O/r02_c7_bm25_block.json 2> $O/r02_c7_bm25_block.err; tail -c 1200 $O/r02_c7_bm25_block.json
timeout 300 python tests/tiny_table_case.py > $O/r02_c7_tiny_table.md 2>&1; cat $O/r02_c7_tiny_table.md
CASE_N=12500000 timeout 300 python tests/batch_case.py 4096,256 1000 1 > $O/r02_c7_c5shard_plain.log 2>&1 &&
CASE_N=12500000 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second \
  --clock-control none -k regex:'select|seed|finalize|override' --csv --log-file $O/r02_c7_c5shard_ncu.csv \
  python tests/batch_case.py 4096,256 1000 1 > $O/r02_c7_c5shard_ncu.log 2>&1
grep -E "finalize|select_tc_kernel<0|seed" $O/r02_c7_c5shard_ncu.csv | grep gpu__time | awk -F'","' '{print $5, $NF}' | head -20
