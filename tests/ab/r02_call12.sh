#!/bin/bash
# round 2, GPU call 12: --set full capture of the fused kernel on bench.py's step (and the B=512 mid-batch case)
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tests/batch_case.py 4096 100 1 > $O/r02_c12_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:select_tc_kernel -s 3 -c 1 \
  -o $O/r02_select_tc_b4096 python tests/batch_case.py 4096 100 1 > $O/r02_c12_ncu_full.log 2>&1
tail -3 $O/r02_c12_ncu_full.log
timeout 300 python tests/batch_case.py 512 100 0 > $O/r02_c12_plain512.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:select_tc_kernel -s 3 -c 1 \
  -o $O/r02_select_tc_b512 python tests/batch_case.py 512 100 0 > $O/r02_c12_ncu_full512.log 2>&1
tail -3 $O/r02_c12_ncu_full512.log
ls -la $O/*.ncu-rep
