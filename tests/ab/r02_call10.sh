#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tests/bm25_bench.py > $O/r02_c10_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_bytes.sum --clock-control none -k regex:'bm25_topk|finalize' --csv --log-file $O/r02_c10_bm25_ncu.csv python tests/bm25_bench.py > $O/r02_c10_ncu.log 2>&1
grep -E "bm25_topk|finalize" $O/r02_c10_bm25_ncu.csv | awk -F'","' '{print $5, $13, $NF}' | head -30
