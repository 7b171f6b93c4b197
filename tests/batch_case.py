#!/usr/bin/env python3
"""One warm-up + one timed public call per batch size over the bench.py corpus, in one process (for
ncu launch lists / DRAM counters per batch size):  python tests/batch_case.py 1,8,128 [k] [mask]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "crowd-coachable-recommendations_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
import ccr_b200  # noqa: E402
from ccr_b200 import engine  # noqa: E402

batches = [int(b) for b in sys.argv[1].split(",")]
k = int(sys.argv[2]) if len(sys.argv) > 2 else bench.TOPK
use_mask = len(sys.argv) > 3 and sys.argv[3] != "0"
n = int(os.environ.get("CASE_N", bench.N_ITEMS))
dev = torch.device("cuda:0")
table = ccr_b200.EmbeddingTable(n, bench.DIM, device=dev)
bench.build_shard(table, 0, n, dev)
for B in batches:
    q = table.encode_queries(torch.randn((B, bench.DIM), generator=torch.Generator().manual_seed(7)))
    mask = None
    if use_mask:
        indptr, cols, vals = bench.rows_to_csr(bench.history_mask_rows(B, n))
        mask = engine.SparseMask(indptr, cols, vals, n, engine.MASK_SET, dev)
    table.search(q, k, mask=mask, encoded=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    table.search(q, k, mask=mask, encoded=True)
    e1.record()
    torch.cuda.synchronize()
    print(f"B={B} k={k} mask={int(use_mask)} ms per call {e0.elapsed_time(e1):.3f}", flush=True)
