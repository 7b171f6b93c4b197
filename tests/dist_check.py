#!/usr/bin/env python3
"""Multi-GPU parity check (run under torchrun on a GPU box; tests/test_gpu_multi.py spawns it):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py
Row-sharded ShardedIndex.search (device mask sharding, local fused top-k, ONE all-gather of packed
keys -- or the (float64, int64) pair exchange under additive priors -- and the G-way merge) must
equal (a) the fp32 device oracle on the whole table under the north_star tolerance rule and (b) the
single-table result of the same kernels bit for bit, on every rank."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "crowd-coachable-recommendations_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import ccr_b200  # noqa: E402
from oracle import ccr_oracle as O  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
ok = True
CASES = [
    # N, D, B, k, mode (0 none, 1 set -1e6, 2 add float64 priors), queries on the host?
    (200_003, 768, 300, 100, 0, True),
    (50_000, 768, 64, 1000, 1, False),
    (3_000, 768, 17, 10, 2, True),
    (1_500, 768, 9, 1001, 1, True),
    (1_200_000, 128, 700, 100, 1, True),    # seeded + histogram + CTA pairs per shard at world <= 4
    (900_000, 128, 130, 1000, 2, False),
    (7, 64, 5, 7, 1, True),                 # fewer rows than ranks: empty shards, k == N
]
for (N, D, B, k, mode, host_q) in CASES:
    g = torch.Generator().manual_seed(N)
    P = torch.randn((N, D), generator=g)
    Q = torch.randn((B, D), generator=g)
    rs = np.random.RandomState(N)
    mask = None
    mhost = None
    if mode:
        rows = [np.unique(rs.randint(0, N, size=rs.randint(0, min(N, 50)))) for _ in range(B)]
        if mode == 1:
            mask = ccr_b200.SparseMask.from_lists(rows, N, -1e6, ccr_b200.MASK_SET, dev)
        else:
            indptr = np.concatenate([[0], np.cumsum([len(r) for r in rows])])
            cols = np.concatenate(rows) if B else np.zeros(0)
            vals = np.where(rs.rand(len(cols)) < 0.5, -1e10, 1e5)
            mask = ccr_b200.SparseMask(indptr, cols, vals, N, ccr_b200.MASK_ADD, dev)
        mhost = mask.host
    full = ccr_b200.EmbeddingTable.from_tensor(P, device=dev)
    s1, i1, d1 = full.search(Q, k, mask=mask, want_f64=True)
    idx = ccr_b200.ShardedIndex(N, D, device=dev)
    idx.add_local(P[idx.lo:idx.hi])
    packed = idx.can_pack(mask)
    assert packed == (mode != 2)
    s2, i2, d2 = idx.search(Q if host_q else Q.to(dev), k, mask=mask)
    same = bool((i1 == i2).all()) and bool((s1 == s2).all())
    if not packed:
        same &= bool((d1 == d2).all())
    # independent arbiter: fp32 (float64 under ADD) torch arithmetic on the encoded operands
    ref_s, ref_i = O.score_topk_ref_device(full.encode_queries(Q), full.rows, k, mask=mhost, mode=mode)
    got = d2 if d2 is not None else s2
    errs = O.check_topk(got.cpu().numpy(), i2.cpu().numpy(), ref_scores=ref_s.numpy(), ref_ids=ref_i.numpy(),
                        rtol=1e-2, atol=1e-4)
    print(f"rank {rank}/{world} N={N} D={D} B={B} k={k} mode={mode} packed={packed} shard=[{idx.lo},{idx.hi}) "
          f"equal_single_table={same} oracle_violations={len(errs)}", flush=True)
    if errs:
        print("   ", errs[:3], flush=True)
    ok &= same and not errs
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
