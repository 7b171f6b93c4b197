#!/usr/bin/env python3
"""One rank's share of the 8-GPU bench step on one GPU: B=4096 queries over a 1,105,228-row shard
(8,841,823 / 8), top-100, history mask (column-sharded), allow_short, float64 scores."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "crowd-coachable-recommendations_b200"))
import numpy as np, torch
import ccr_b200
from ccr_b200 import engine
N, B, k = 8_841_823, 4096, 100
lo, hi = ccr_b200.shard_bounds(N, 8, 3)
dev = torch.device("cuda:0")
table = ccr_b200.EmbeddingTable(hi - lo, 768, device=dev, id_offset=lo)
g = torch.Generator(device=dev).manual_seed(1)
table.append(torch.randn((hi - lo, 768), generator=g, device=dev))
rs = np.random.RandomState(2)
rows = [np.unique(rs.randint(0, N, size=min(64, rs.geometric(1.0 / 8)))) for _ in range(B)]
mask = engine.SparseMask.from_lists(rows, N, -1e6, engine.MASK_SET, dev).column_shard(lo, hi)
q = table.encode_queries(torch.randn((B, 768), generator=torch.Generator().manual_seed(7)))
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for _ in range(2):
    table.search(q, k, mask=mask, allow_short=True, want_f64=True, encoded=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    s, i, d = table.search(q, k, mask=mask, allow_short=True, want_f64=True, encoded=True)
    gs = d.unsqueeze(0).repeat(8, 1, 1); gi = i.unsqueeze(0).repeat(8, 1, 1)
    engine.merge_topk(gs, gi, k)
e1.record(); torch.cuda.synchronize()
print("ms per step (local + 8-way merge, no NCCL)", e0.elapsed_time(e1) / iters)
