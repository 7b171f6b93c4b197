#!/usr/bin/env python3
"""BASELINE.json configs[3]: query-batch sweep 1..16384 over the 8.84M x 768 corpus (and the NQ shape),
top-100, no mask.  Prints a markdown table: ms, queries/s, fraction of the binding roofline
(min of tensor peak and HBM bandwidth over the corpus bytes, MEASURED_PEAKS burst figures)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "crowd-coachable-recommendations_b200"))
import torch  # noqa: E402

from ccr_b200 import engine  # noqa: E402

pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
F, BW = pk.get("bf16_tflops", 1667.8) * 1e12, pk.get("hbm_gbs", 6445.3) * 1e9
dev = torch.device("cuda:0")


def table(n):
    t = torch.empty((n, 768), dtype=torch.bfloat16, device=dev)
    g = torch.Generator(device=dev).manual_seed(1)
    for s in range(0, n, 1 << 20):
        e = min(n, s + (1 << 20))
        t[s:e] = torch.randn((e - s, 768), generator=g, device=dev).to(torch.bfloat16)
    return t


def run(items, B, k=100, iters=None):
    N = items.shape[0]
    q = torch.randn((B, 768), generator=torch.Generator(device=dev).manual_seed(7), device=dev).to(torch.bfloat16)
    iters = iters or max(3, min(20, int(0.3 / max(2e-3, 13e-6 * B * N / 8.8e6))))
    for _ in range(2):
        engine.score_topk(q, items, k)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        engine.score_topk(q, items, k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    t_min = max(2.0 * B * N * 768 / F, N * 768 * 2 / BW)
    return ms, B / ms * 1e3, t_min * 1e3 / ms


print("| corpus | B | ms/batch | queries/s | bound | frac of roofline |")
print("|---|---|---|---|---|---|")
only = [int(x) for x in sys.argv[1:]]  # optional: restrict the MS-MARCO sweep to these batch sizes
for name, n, batches in (("MS-MARCO 8,841,823", 8841823, only or [1, 2, 4, 8, 16, 32, 64, 128, 256, 384, 512, 1024, 2048, 4096, 8192, 16384]),
                         ("NQ 2,681,468", 2681468, [] if only else [128, 512, 3452])):
    items = table(n)
    for B in batches:
        ms, qps, frac = run(items, B)
        bound = "tensor" if 2.0 * B * 768 / F > 768 * 2 / BW else "hbm"
        print(f"| {name} | {B} | {ms:.3f} | {qps:,.0f} | {bound} | {frac:.3f} |", flush=True)
    del items
    torch.cuda.empty_cache()
