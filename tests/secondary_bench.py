#!/usr/bin/env python3
"""Throughput of the secondary device paths against their rooflines (builder tool, one JSON line each):

* ``engine.topk_dense``     -- top-k of an already materialised float32 score matrix (`_assign_topk` on the
                               reference's unmodified ``transform`` output): HBM-bound, 4 bytes per score;
* ``engine.argsort_scores`` -- `_argsort`: LSD radix sort of the whole matrix, 8 passes x 24 bytes per score;
* ``table.dense_scores``    -- `MatMulExpression.as_tensor("cuda")`: tcgen05 store-mode tiles, 4 bytes written per score.

    python tests/secondary_bench.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "crowd-coachable-recommendations_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
import ccr_b200  # noqa: E402
from ccr_b200 import engine  # noqa: E402

dev = torch.device("cuda:0")
HBM = bench.peaks()["hbm"]


def timed(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


g = torch.Generator(device=dev).manual_seed(5)
for B, N, k in ((4096, 262144, 100), (512, 2097152, 100), (4096, 262144, 1001)):
    S = torch.randn((B, N), generator=g, device=dev)
    ms = timed(lambda: engine.topk_dense(S, k))
    gbs = B * N * 4 / ms / 1e6
    print(json.dumps({"what": "topk_dense (ccr_topk_dense_f32)", "B": B, "N": N, "k": k, "ms": ms, "GBps": gbs,
                      "frac_of_hbm": gbs / HBM}), flush=True)
    del S

for B, N in ((256, 262144), (64, 1048576)):
    S = torch.randn((B, N), generator=g, device=dev)
    ms = timed(lambda: engine.argsort_scores(S, None), iters=3)
    n = B * N
    print(json.dumps({"what": "argsort_scores (ccr_argsort_scores_f32)", "B": B, "N": N, "ms": ms,
                      "Mkeys_per_s": n / ms / 1e3, "GBps_at_8x24B": n * 8 * 24 / ms / 1e6,
                      "frac_of_hbm_at_8x24B": n * 8 * 24 / ms / 1e6 / HBM}), flush=True)
    tms = timed(lambda: torch.sort(S.flatten(), descending=True, stable=True), iters=3)
    print(json.dumps({"what": "torch.sort of the same flat matrix (library)", "B": B, "N": N, "ms": tms}), flush=True)
    del S

table = ccr_b200.EmbeddingTable(1 << 20, bench.DIM, device=dev)
bench.build_shard(table, 0, 1 << 20, dev)
for B in (512, 2048):
    q = torch.randn((B, bench.DIM), generator=torch.Generator().manual_seed(7))
    ms = timed(lambda: table.dense_scores(q), iters=3)
    n = B * (1 << 20)
    print(json.dumps({"what": "dense_scores (ccr_score_dense_f32, tcgen05 store tiles)", "B": B, "N": 1 << 20, "ms": ms,
                      "TFLOPs": 2.0 * n * bench.DIM / ms / 1e9, "write_GBps": n * 4 / ms / 1e6,
                      "frac_of_hbm_write": n * 4 / ms / 1e6 / HBM}), flush=True)
