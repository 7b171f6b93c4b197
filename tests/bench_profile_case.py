#!/usr/bin/env python3
"""The bench.py workload (B=4096, 8.84M x 768, top-100, history mask) as ONE call, for ncu."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "crowd-coachable-recommendations_b200"))
import torch  # noqa: E402

import ccr_b200  # noqa: E402
from ccr_b200 import engine  # noqa: E402

import numpy as np  # noqa: E402


class bench:  # the constants / generators of bench.py, restated so this script has no other imports
    N_ITEMS, DIM, TOPK, CHUNK = 8_841_823, 768, 100, 1 << 20

    @staticmethod
    def build_shard(table, lo, hi, dev):
        for c in range(lo // bench.CHUNK, (hi + bench.CHUNK - 1) // bench.CHUNK):
            g = torch.Generator(device=dev).manual_seed(1000 + c)
            rows = torch.randn((bench.CHUNK, bench.DIM), generator=g, device=dev)
            a, b = max(lo, c * bench.CHUNK), min(hi, (c + 1) * bench.CHUNK)
            table.append(rows[a - c * bench.CHUNK : b - c * bench.CHUNK])

    @staticmethod
    def history_mask_rows(B, n_items, seed=2):
        rs = np.random.RandomState(seed)
        return [np.unique(rs.randint(0, n_items, size=min(64, rs.geometric(1.0 / 8)))) for _ in range(B)]

    @staticmethod
    def rows_to_csr(rows):
        indptr = np.zeros(len(rows) + 1, dtype=np.int64)
        np.cumsum([len(r) for r in rows], out=indptr[1:])
        cols = np.concatenate(rows).astype(np.int32)
        return indptr, cols, np.full(len(cols), -1e6)


B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1
if len(sys.argv) > 3:
    bench.N_ITEMS = int(sys.argv[3])  # e.g. 1105228 = one rank's shard at 8 GPUs
dev = torch.device("cuda:0")
table = ccr_b200.EmbeddingTable(bench.N_ITEMS, bench.DIM, device=dev)
bench.build_shard(table, 0, bench.N_ITEMS, dev)
q = table.encode_queries(torch.randn((B, bench.DIM), generator=torch.Generator().manual_seed(7)))
indptr, cols, vals = bench.rows_to_csr(bench.history_mask_rows(B, bench.N_ITEMS))
mask = engine.SparseMask(indptr, cols, vals, bench.N_ITEMS, engine.MASK_SET, dev)
table.search(q, bench.TOPK, mask=mask, encoded=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    table.search(q, bench.TOPK, mask=mask, encoded=True)
e1.record()
torch.cuda.synchronize()
print("ms per step", e0.elapsed_time(e1) / iters)
