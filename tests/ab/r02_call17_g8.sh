#!/bin/bash
# round 2, GPU call 17 (8 GPUs, final build): sharded search vs oracle at world 2/4/8, bench.py at 8 and 4 ranks, configs[4]
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/r02_c17_multi.log 2>&1; echo "pytest rc $?" >> $O/r02_c17_multi.log; tail -8 $O/r02_c17_multi.log
for G in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2954$G \
  bench.py --gpus $G --steps 20 --warmup 3 > $O/r02_c17_bench_g$G.json 2> $O/r02_c17_bench_g$G.err; tail -c 1800 $O/r02_c17_bench_g$G.json; tail -2 $O/r02_c17_bench_g$G.err
done
for B in 4096 256; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2956$((B/256%10)) \
  bench.py --gpus 8 --config c5 --batch $B --steps 10 --warmup 3 > $O/r02_c17_c5_b$B.json 2> $O/r02_c17_c5_b$B.err; tail -c 1800 $O/r02_c17_c5_b$B.json; tail -2 $O/r02_c17_c5_b$B.err
done
