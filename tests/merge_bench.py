#!/usr/bin/env python3
"""Time ccr_merge_topk at the exchange shape of the 100M / 8-GPU config: G=8 runs of k=1000 for
B=4096 rows (and G=8, k=100).  python tests/merge_bench.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "crowd-coachable-recommendations_b200"))
import torch  # noqa: E402

from ccr_b200 import engine  # noqa: E402

dev = torch.device("cuda:0")
for G, B, k in ((8, 4096, 1000), (8, 4096, 100), (2, 4096, 100)):
    g = torch.Generator(device=dev).manual_seed(1)
    sc = torch.sort(torch.randn((G, B, k), generator=g, device=dev, dtype=torch.float64), dim=2, descending=True).values
    ids = torch.randperm(G * B * k, generator=g, device=dev).reshape(G, B, k)
    engine.merge_topk(sc, ids, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        s, i, d = engine.merge_topk(sc, ids, k)
    e1.record()
    torch.cuda.synchronize()
    flat_s = sc.permute(1, 0, 2).reshape(B, G * k)
    want = torch.sort(flat_s, dim=1, descending=True, stable=True).values[:, :k]
    print(f"G={G} B={B} k={k}: {e0.elapsed_time(e1) / 10:.3f} ms per merge, exact={bool(torch.equal(d, want))}", flush=True)
