#!/usr/bin/env python3
"""BASELINE.json configs other than the bench line, timed on one B200 (device time, CUDA events):
  C2  NQ shape 2,681,468 x 768, B=3,452, top-100 and top-1001 (what ranking() really asks for)
  C5s one shard of the 100M x 768 / 8-GPU config: 12,500,000 x 768, top-1000, B=256 and 4,096
Prints markdown rows: config | B | k | ms | queries/s | frac of binding (burst) roofline."""
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


def run(name, items, B, k):
    N = items.shape[0]
    q = torch.randn((B, 768), generator=torch.Generator(device=dev).manual_seed(7), device=dev).to(torch.bfloat16)
    for _ in range(2):
        engine.score_topk(q, items, k)
    iters = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        s, i = engine.score_topk(q, items, k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    # spot check: returned scores are the items' own scores, rows sorted, ids unique
    r = 0
    ref = (q[r : r + 1].float() @ items[i[r]].float().T)[0]
    ok = bool(torch.allclose(ref, s[r], rtol=1e-2, atol=1e-3)) and bool((s[r][1:] <= s[r][:-1]).all()) and \
        int(torch.unique(i[r]).numel()) == k
    t_min = max(2.0 * B * N * 768 / F, N * 768 * 2 / BW)
    print(f"| {name} | {B} | {k} | {ms:.3f} | {B / ms * 1e3:,.0f} | {t_min * 1e3 / ms:.3f} | {'ok' if ok else 'BAD'} |", flush=True)


print("| config | B | k | ms/batch | queries/s | frac of roofline | spot check |")
print("|---|---|---|---|---|---|---|")
which = sys.argv[1:] or ["c2", "c5s"]
if "c2" in which:
    items = table(2_681_468)
    for k in (100, 1001):
        run("C2 NQ 2,681,468", items, 3452, k)
    del items
    torch.cuda.empty_cache()
if "c5s" in which:
    items = table(12_500_000)
    for B in (256, 4096):
        run("C5 shard 12,500,000", items, B, 1000)
    run("C5 shard 12,500,000", items, 4096, 100)
