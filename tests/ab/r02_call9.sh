#!/bin/bash
# round 2, GPU call 9: BM25 doc-split granularity (load balance across Zipf queries)
mkdir -p gpurun_out
O=gpurun_out
for W in 1 8 16 32 64; do
CCR_BM25_WAVES=$W timeout 600 python tests/bm25_bench.py 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('waves $W', round(d['ms_per_batch'],2), 'ms', round(d['postings_GBps']), 'GB/s', d['mismatches_vs_torch_f64'])" | tee -a $O/r02_c9_bm25_waves.txt
done
