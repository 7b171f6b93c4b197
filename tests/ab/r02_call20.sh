#!/bin/bash
# round 2, GPU call 20 (final BM25 build): parity tests, head-fraction sweep at the NQ shape
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_bm25.py -m gpu -q -x > $O/r02_c20_pytest.log 2>&1; echo "pytest rc $?" >> $O/r02_c20_pytest.log
tail -4 $O/r02_c20_pytest.log
BM25_KERNELS=${BM25_KERNELS:-auto} BM25_HEAD_FRAC=${BM25_HEAD_FRAC:-0,0.67,0.5,0.33,0.25,0.15} timeout 600 python tests/bm25_bench.py > $O/r02_c20_bm25.jsonl 2> $O/r02_c20_bm25.err
python - <<'PY'
import json
for l in open("gpurun_out/r02_c20_bm25.jsonl"):
    d = json.loads(l)
    print(d["kernel"], d["head_df_fraction"], d["head_terms"], "%.2f ms" % d["ms_per_batch"], "%.0f q/s" % d["queries_per_s"],
          "%.0f GB/s postings" % d["postings_GBps"], "%.0f GB/s moved" % d["bytes_moved_GBps"], "bad", d["mismatches_vs_torch_f64"])
PY
tail -3 $O/r02_c20_bm25.err
