#!/bin/bash
# round 2, GPU call 23 (2 GPUs): multi-GPU parity tests + bench.py under torchrun on the final build
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/r02_c23_pytest.log 2>&1; echo "pytest rc $?" >> $O/r02_c23_pytest.log
tail -4 $O/r02_c23_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --steps 10 --warmup 3 > $O/r02_c23_bench_g2.json 2> $O/r02_c23_bench_g2.err; tail -c 2500 $O/r02_c23_bench_g2.json; tail -3 $O/r02_c23_bench_g2.err
