#!/usr/bin/env python3
"""One fused call with CCR_DEBUG=128 (device printf timeline of block 0 / warp 0): how fast the
thresholds mature.  python tests/timeline_case.py [B] [k] [N]"""
import os
import sys

os.environ.setdefault("CCR_DEBUG", "128")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "crowd-coachable-recommendations_b200"))
import torch  # noqa: E402

from ccr_b200 import engine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
k = int(sys.argv[2]) if len(sys.argv) > 2 else 100
N = int(sys.argv[3]) if len(sys.argv) > 3 else 8_841_823
dev = torch.device("cuda:0")
t = torch.empty((N, 768), dtype=torch.bfloat16, device=dev)
g = torch.Generator(device=dev).manual_seed(1)
for s in range(0, N, 1 << 20):
    e = min(N, s + (1 << 20))
    t[s:e] = torch.randn((e - s, 768), generator=g, device=dev).to(torch.bfloat16)
q = torch.randn((B, 768), generator=torch.Generator(device=dev).manual_seed(7), device=dev).to(torch.bfloat16)
engine.score_topk(q, t, k)
torch.cuda.synchronize()
