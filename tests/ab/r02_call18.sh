#!/bin/bash
# round 2, GPU call 18: what would an on-chip-resident query operand buy?  CCR_DEBUG=1024 streams the query
# block only for the first tile of a unit (results invalid, timing meaningful); =1 drops the selection work.
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python tests/perf_sweep.py --batches 256,512,1024,4096 \
  --variants "base=;noq=CCR_DEBUG=1024;gemm=CCR_DEBUG=1;gemm_noq=CCR_DEBUG=1025" --secs 0.5 --rounds 2 \
  --md $O/r02_c18_qresident.md > $O/r02_c18_qresident.log 2>&1; tail -18 $O/r02_c18_qresident.log
