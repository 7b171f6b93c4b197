#!/bin/bash
# round 2, GPU call 11 (2 GPUs): query-sharded exchange vs oracle, bench on 2 ranks, BM25 after the finalize bound
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_bm25.py -m gpu -q > $O/r02_c11_pytest.log 2>&1; echo "pytest rc $?" >> $O/r02_c11_pytest.log; tail -6 $O/r02_c11_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 2 --steps 10 --warmup 3 > $O/r02_c11_bench_g2.json 2> $O/r02_c11_bench_g2.err; tail -c 1500 $O/r02_c11_bench_g2.json; tail -3 $O/r02_c11_bench_g2.err
timeout 600 python tests/bm25_bench.py 2>/dev/null | tail -1 | tee $O/r02_c11_bm25.json
