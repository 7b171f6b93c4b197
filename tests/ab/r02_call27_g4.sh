#!/bin/bash
# round 2, GPU call 27 (4 GPUs): multi-GPU parity tests (2 and 4 ranks) + bench.py under torchrun on the last build
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/r02_c27_pytest.log 2>&1; echo "pytest rc $?" >> $O/r02_c27_pytest.log
tail -4 $O/r02_c27_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 \
  bench.py --gpus 4 --steps 10 --warmup 3 > $O/r02_c27_bench_g4.json 2> $O/r02_c27_bench_g4.err; tail -c 900 $O/r02_c27_bench_g4.json; tail -2 $O/r02_c27_bench_g4.err
