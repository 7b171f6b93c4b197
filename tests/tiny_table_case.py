#!/usr/bin/env python3
"""Tiny tables with a handful of query rows: the CUDA-core streaming kernel (CCR_ALGO_SIMT) against the
TMA + tcgen05 kernel, microseconds per public call (back-to-back, CUDA events).  Decides what
ccr_choose_algo does below 65,536 rows."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "crowd-coachable-recommendations_b200"))
import torch  # noqa: E402

from ccr_b200 import engine  # noqa: E402

dev = torch.device("cuda:0")
print("| N | B | k | SIMT us/call | tcgen05 us/call |")
print("|---|---|---|---|---|")
for N in (1000, 9862, 20000, 65535, 262144):
    items = torch.randn((N, 768), device=dev).to(torch.bfloat16)
    for B in (1, 8):
        for k in (10, 100):
            q = torch.randn((B, 768), device=dev).to(torch.bfloat16)
            out = []
            for algo in (1, 2):
                for _ in range(10):
                    engine.score_topk(q, items, k, algo=algo)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(200):
                    engine.score_topk(q, items, k, algo=algo)
                e1.record()
                torch.cuda.synchronize()
                out.append(e0.elapsed_time(e1) / 200 * 1e3)
            print(f"| {N} | {B} | {k} | {out[0]:.1f} | {out[1]:.1f} |", flush=True)
