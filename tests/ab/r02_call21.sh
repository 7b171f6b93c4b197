#!/bin/bash
# round 2, GPU call 21: ncu --set full with source counters of the BM25 warp-private kernel (512 queries, head rows)
mkdir -p gpurun_out
O=gpurun_out
BM25_HEAD_FRAC=0.25 timeout 300 python tests/bm25_bench.py 2681468 512 1001 > $O/r02_c21_plain.jsonl 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bm25_topk_warp -s 2 -c 1 \
  -o $O/r02_bm25_warp python tests/bm25_bench.py 2681468 512 1001 > $O/r02_c21_ncu.log 2>&1
tail -2 $O/r02_c21_ncu.log; cut -c 1-300 $O/r02_c21_plain.jsonl
