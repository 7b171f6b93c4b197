#!/bin/bash
# round 2, GPU call 8: BM25 warp kernel v2 vs block-wide, BM25 tests
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_bm25.py tests/test_gpu_fullsize.py -m gpu -q --maxfail=15 > $O/r02_c8_pytest.log 2>&1; echo "pytest rc $?" >> $O/r02_c8_pytest.log
tail -5 $O/r02_c8_pytest.log
timeout 600 python tests/bm25_bench.py > $O/r02_c8_bm25_warp.json 2> $O/r02_c8_bm25_warp.err; tail -c 700 $O/r02_c8_bm25_warp.json; tail -2 $O/r02_c8_bm25_warp.err
BM25_KERNELS=blockwide timeout 600 python tests/bm25_bench.py > $O/r02_c8_bm25_block.json 2> $O/r02_c8_bm25_block.err; tail -c 700 $O/r02_c8_bm25_block.json
