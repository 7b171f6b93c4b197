#!/usr/bin/env python3
"""Multi-GPU parity check (run under torchrun on a GPU box):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py
Row-sharded ShardedIndex.search (local fused top-k + all-gather + G-way merge) must equal the
single-table result computed on every rank from the same seeded data."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "crowd-coachable-recommendations_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import ccr_b200  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
ok = True
for (N, B, k, mode) in [(200_003, 300, 100, 0), (50_000, 64, 1000, 1), (3_000, 17, 10, 2), (1_500, 9, 1001, 1)]:
    g = torch.Generator().manual_seed(N)
    P = torch.randn((N, 768), generator=g)
    Q = torch.randn((B, 768), generator=g)
    rs = np.random.RandomState(N)
    mask = None
    if mode:
        rows = [np.unique(rs.randint(0, N, size=rs.randint(0, 50))) for _ in range(B)]
        if mode == 1:
            mask = ccr_b200.SparseMask.from_lists(rows, N, -1e6, ccr_b200.MASK_SET, dev)
        else:
            indptr = np.concatenate([[0], np.cumsum([len(r) for r in rows])])
            cols = np.concatenate(rows)
            vals = np.where(rs.rand(len(cols)) < 0.5, -1e10, 1e5)
            mask = ccr_b200.SparseMask(indptr, cols, vals, N, ccr_b200.MASK_ADD, dev)
    full = ccr_b200.EmbeddingTable.from_tensor(P, device=dev)
    s1, i1, d1 = full.search(Q, k, mask=mask, want_f64=True)
    idx = ccr_b200.ShardedIndex(N, 768, device=dev)
    idx.add_local(P[idx.lo:idx.hi])
    s2, i2, d2 = idx.search(Q, k, mask=mask)
    same = bool((i1 == i2).all()) and bool((d1 == d2).all())
    print(f"rank {rank}/{world} N={N} B={B} k={k} mode={mode} shard=[{idx.lo},{idx.hi}) equal={same}", flush=True)
    ok &= same
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
