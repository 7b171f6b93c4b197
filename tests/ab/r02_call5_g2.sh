#!/bin/bash
# round 2, GPU call 5 (2 GPUs): sharded search vs oracle at world 2, bench.py on 2 ranks
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/r02_c5_multi.log 2>&1; echo "pytest rc $?" >> $O/r02_c5_multi.log; tail -15 $O/r02_c5_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 2 --steps 10 --warmup 3 > $O/r02_c5_bench_g2.json 2> $O/r02_c5_bench_g2.err; tail -c 2500 $O/r02_c5_bench_g2.json; tail -3 $O/r02_c5_bench_g2.err
