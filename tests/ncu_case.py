#!/usr/bin/env python3
"""One fused call for profiling:  python tests/ncu_case.py B N k algo [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "crowd-coachable-recommendations_b200"))
import torch  # noqa: E402

from ccr_b200 import engine  # noqa: E402

B, N, k, algo = (int(x) for x in sys.argv[1:5])
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 1
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
items = torch.empty((N, 768), dtype=torch.bfloat16, device=dev)
for s in range(0, N, 1 << 20):
    e = min(N, s + (1 << 20))
    items[s:e] = torch.randn((e - s, 768), generator=g, device=dev).to(torch.bfloat16)
q = torch.randn((B, 768), generator=g, device=dev).to(torch.bfloat16)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
engine.score_topk(q, items, k, algo=algo)
torch.cuda.synchronize()
ev0.record()
for _ in range(iters):
    engine.score_topk(q, items, k, algo=algo)
ev1.record()
torch.cuda.synchronize()
print("ms per call", ev0.elapsed_time(ev1) / iters)
