#!/usr/bin/env python3
"""Context number: the "library-call" GPU implementation of the same step -- stock
torch bf16 ``Q @ V_chunk.T`` (cuBLAS) + ``torch.topk`` per chunk + a final merge -- on the bench
workload shape (B=4096, 8,841,823 x 768, top-100, no mask), next to the fused kernel on the same box.
The reference ships no GPU kernel of its own; this is what a torch user would write."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "crowd-coachable-recommendations_b200"))
import torch  # noqa: E402

from ccr_b200 import engine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N, K, CH = 8_841_823, 100, 1 << 17
dev = torch.device("cuda:0")
items = torch.empty((N, 768), dtype=torch.bfloat16, device=dev)
g = torch.Generator(device=dev).manual_seed(1)
for s in range(0, N, 1 << 20):
    e = min(N, s + (1 << 20))
    items[s:e] = torch.randn((e - s, 768), generator=g, device=dev).to(torch.bfloat16)
q = torch.randn((B, 768), generator=torch.Generator(device=dev).manual_seed(7), device=dev).to(torch.bfloat16)


def torch_step():
    best_s, best_i = [], []
    for s in range(0, N, CH):
        e = min(N, s + CH)
        sc = q @ items[s:e].T  # bf16 out, fp32 accumulate (cuBLAS)
        v, i = sc.topk(K, dim=1)
        best_s.append(v)
        best_i.append(i + s)
    v = torch.cat(best_s, 1)
    i = torch.cat(best_i, 1)
    top, pos = v.topk(K, dim=1)
    return top, torch.gather(i, 1, pos)


def timeit(fn, iters):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, out


ms_t, (ts, ti) = timeit(torch_step, 3)
ms_o, (os_, oi) = timeit(lambda: engine.score_topk(q, items, K), 5)
# agreement of the two id sets (torch ranks bf16-rounded scores, so near-ties may differ)
same = float((torch.sort(ti, 1).values == torch.sort(oi, 1).values).float().mean())
print(json.dumps({"B": B, "n_items": N, "k": K, "torch_library_ms": ms_t, "torch_library_qps": B / ms_t * 1e3,
                  "fused_ms": ms_o, "fused_qps": B / ms_o * 1e3, "speedup": ms_t / ms_o,
                  "id_agreement_sorted_positions": same}))
