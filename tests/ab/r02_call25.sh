#!/bin/bash
# round 2, GPU call 25: ncu --set full with source counters of the fused kernel on the bench step (B=4096, mask)
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tests/batch_case.py 4096 100 1 > $O/r02_c25_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:select_tc_kernel -s 3 -c 1 \
  -o $O/r02_select_tc_b4096_src python tests/batch_case.py 4096 100 1 > $O/r02_c25_ncu.log 2>&1
tail -2 $O/r02_c25_ncu.log; cat $O/r02_c25_plain.log | tail -1
