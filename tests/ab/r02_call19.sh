#!/bin/bash
# round 2, GPU call 19: BM25 hybrid index (dense float64 rows for head terms): parity tests, head-fraction sweep
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_bm25.py -m gpu -q -x > $O/r02_c19_pytest.log 2>&1; echo "pytest rc $?" >> $O/r02_c19_pytest.log
tail -5 $O/r02_c19_pytest.log
BM25_HEAD_FRAC=0,0.67,0.5,0.33,0.25,0.15,0.08 timeout 600 python tests/bm25_bench.py > $O/r02_c19_bm25.jsonl 2> $O/r02_c19_bm25.err
python - <<'PY'
import json
for l in open("gpurun_out/r02_c19_bm25.jsonl"):
    d = json.loads(l)
    print(d["head_df_fraction"], d["head_terms"], "%.2f ms" % d["ms_per_batch"], "%.0f q/s" % d["queries_per_s"],
          "%.0f GB/s postings" % d["postings_GBps"], "%.0f GB/s moved" % d["bytes_moved_GBps"], "bad", d["mismatches_vs_torch_f64"])
PY
tail -3 $O/r02_c19_bm25.err
CCR_B200_LIB=$PWD/crowd-coachable-recommendations_b200/lib/libccr_b200_bmw4.so BM25_HEAD_FRAC=0,0.25 timeout 600 python tests/bm25_bench.py > $O/r02_c19_bm25_bmw4.jsonl 2> $O/r02_c19_bm25_bmw4.err
cut -c 1-400 $O/r02_c19_bm25_bmw4.jsonl | grep -o '"ms_per_batch": [0-9.]*\|"head_df_fraction": [0-9.]*'
