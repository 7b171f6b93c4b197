#!/bin/bash
# round 2, GPU call 27: BM25: trimmed range-streaming kernel vs the cursor kernels (5 and 4 blocks per SM)
mkdir -p gpurun_out
O=gpurun_out
D=$PWD/crowd-coachable-recommendations_b200/lib
timeout 900 python -m pytest tests/test_gpu_bm25.py -m gpu -q -x > $O/r02_c27_pytest.log 2>&1; echo "pytest rc $?" >> $O/r02_c27_pytest.log
tail -4 $O/r02_c27_pytest.log
BM25_LIBS=$D/libccr_b200.so,$D/libccr_b200_bmw4.so BM25_KERNELS=stream,cursor \
BM25_HEAD_FRAC=0,0.25 timeout 600 python tests/bm25_bench.py > $O/r02_c27_bm25.jsonl 2> $O/r02_c27_bm25.err
python - <<'PY'
import json
for l in open("gpurun_out/r02_c27_bm25.jsonl"):
    d = json.loads(l)
    print(d["kernel"], d["head_df_fraction"], "%.2f ms" % d["ms_per_batch"], "bad", d["mismatches_vs_torch_f64"])
PY
tail -3 $O/r02_c27_bm25.err
