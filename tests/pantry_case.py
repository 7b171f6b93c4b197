#!/usr/bin/env python3
"""BASELINE.json configs[0] at its real shape (Prime Pantry: 9,862 items = queries = corpus, 768-d,
1,960 brands, same-brand block mask incl. the query itself, top-1001 per query), synthetic
embeddings: the full ``ranking()`` drop-in on the GPU next to the oracle's CPU
restatement of the reference, with the parity check over every row.  Prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "crowd-coachable-recommendations_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import cases  # noqa: E402
import ccr_b200  # noqa: E402
from oracle import ccr_oracle as O  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 9862
os.environ["CCREC_SIM_TYPE"] = "dot"
emb = torch.randn((N, 768), generator=torch.Generator().manual_seed(0))
corpus = {f"p{i}": f"text#{i}" for i in range(N)}
brand = np.random.RandomState(1).zipf(1.1, size=N) % 1960
groups = {}
for i, b in enumerate(brand):
    groups.setdefault(int(b), []).append(f"p{i}")
block_dict = {f"p{i}": groups[int(brand[i])] for i in range(N)}

table = cases.TextTable(emb)
ccr_b200.ranking(dict(list(corpus.items())[:64]), dict(list(corpus.items())[:8]), table, 512,
                 {q: [q] for q in list(corpus)[:8]})  # warm-up (module load, workspace)
torch.cuda.synchronize()
t0 = time.perf_counter()
prof = ccr_b200.ranking(corpus, corpus, table, 512, block_dict)
torch.cuda.synchronize()
t_gpu = time.perf_counter() - t0

t0 = time.perf_counter()
want = O.ranking_ref(corpus, corpus, cases.TextTable(emb), 512, block_dict, sim_type="dot")
t_cpu = time.perf_counter() - t0

pos = {p: i for i, p in enumerate(corpus)}
gs = np.array([list(prof[q].values()) for q in corpus])
gi = np.array([[pos[p] for p in prof[q]] for q in corpus])
ws = np.array([list(want[q].values()) for q in corpus])
wi = np.array([[pos[p] for p in want[q]] for q in corpus])
errs = O.check_topk(gs, gi, ref_scores=ws, ref_ids=wi, rtol=1e-2, atol=1e-3)
top2_same = float(np.mean((gi[:, :2] == wi[:, :2]).all(1)))
top100_overlap = float(np.mean([len(set(a[:100]) & set(b[:100])) / 100 for a, b in zip(gi, wi)]))
print(json.dumps({"config": "prime-pantry shape", "n_items": N, "n_queries": N, "k": gs.shape[1], "brands": 1960,
                  "largest_block": int(max(len(v) for v in groups.values())),
                  "ranking_gpu_s": t_gpu, "ranking_cpu_oracle_s": t_cpu, "cpu_threads": torch.get_num_threads(),
                  "speedup_whole_call": t_cpu / t_gpu, "tolerance_rule_violations": len(errs),
                  "first_violations": errs[:3], "ordered_top2_identical_rows": top2_same,
                  "top100_set_overlap": top100_overlap}))
