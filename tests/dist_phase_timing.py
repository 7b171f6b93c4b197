#!/usr/bin/env python3
"""Where does a sharded step spend its time?  Run under torchrun on G GPUs; every rank holds
`rows_per_rank` items (default 1,105,228 = the 8-GPU shard of the 8.84M corpus), B=4096, k=100,
history mask.  CUDA events around: local fused top-k | all-gathers | merge.  Prints rank 0's medians."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "crowd-coachable-recommendations_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import ccr_b200  # noqa: E402
from ccr_b200 import engine  # noqa: E402
import bench  # noqa: E402

rows_per_rank = int(sys.argv[1]) if len(sys.argv) > 1 else 1_105_228
B, k = 4096, 100
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
N = rows_per_rank * world
index = ccr_b200.ShardedIndex(N, 768, device=dev)
bench.build_shard(index.table, index.lo, index.hi, dev)
q = index.table.encode_queries(torch.randn((B, 768), generator=torch.Generator().manual_seed(7)))
indptr, cols, vals = bench.rows_to_csr(bench.history_mask_rows(B, N))
mask = engine.SparseMask(indptr, cols, vals, N, engine.MASK_SET, dev).column_shard(index.lo, index.hi)


def step(ev):
    ev[0].record()
    d, i = index._local_topk(q, k, mask)
    ev[1].record()
    gs = torch.empty((world * B, k), dtype=d.dtype, device=dev)
    gi = torch.empty((world * B, k), dtype=i.dtype, device=dev)
    dist.all_gather_into_tensor(gs, d)
    dist.all_gather_into_tensor(gi, i)
    ev[2].record()
    out = engine.merge_topk(gs.view(world, B, k), gi.view(world, B, k), k)
    ev[3].record()
    return out


mk = lambda: [torch.cuda.Event(enable_timing=True) for _ in range(4)]  # noqa: E731
for _ in range(3):
    step(mk())
dist.barrier()
torch.cuda.synchronize()
evs = [mk() for _ in range(10)]
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for e in evs:
    step(e)
t1.record()
torch.cuda.synchronize()
if rank == 0:
    med = lambda a: float(np.median(a))  # noqa: E731
    print({"world": world, "rows_per_rank": rows_per_rank, "ms_per_step": t0.elapsed_time(t1) / 10,
           "local_topk_ms": med([e[0].elapsed_time(e[1]) for e in evs]),
           "allgather_ms": med([e[1].elapsed_time(e[2]) for e in evs]),
           "merge_ms": med([e[2].elapsed_time(e[3]) for e in evs])}, flush=True)
dist.barrier()
dist.destroy_process_group()
