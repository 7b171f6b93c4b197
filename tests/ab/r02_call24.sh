#!/bin/bash
# round 2, GPU call 24: ncu launch list of the BM25 bench (kernel vs finalize share), full batch
mkdir -p gpurun_out
O=gpurun_out
BM25_HEAD_FRAC=0.25 timeout 300 python tests/bm25_bench.py > $O/r02_c24_plain.jsonl 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_bytes.sum --clock-control none -k regex:'bm25_topk|finalize' -c 12 --csv \
  --log-file $O/r02_bm25_launches.csv python tests/bm25_bench.py > $O/r02_c24_ncu.log 2>&1
tail -2 $O/r02_c24_ncu.log; cut -c 1-200 $O/r02_c24_plain.jsonl
