#!/usr/bin/env python3
"""BM25 kernel throughput at the NQ shape (the reference needs 21 min 19 s for 3,452
queries x 2.68 M passages on CPU).  Synthetic postings built on the device (log-uniform ~ Zipf(1)
term frequencies, ~40 tokens per doc), scored through the C ABI; a few rows are cross-checked
against a plain torch float64 recomputation.  Prints one JSON line.

    python tests/bm25_bench.py [n_docs] [n_queries] [k]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "crowd-coachable-recommendations_b200"))
import torch  # noqa: E402

from ccr_b200 import _lib  # noqa: E402
from ccr_b200.engine import _stream_ptr  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2_681_468
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 3452
K = int(sys.argv[3]) if len(sys.argv) > 3 else 1001
V, TOK = 200_000, 40
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(3)


def zipf_terms(n):
    u = torch.rand(n, generator=g, device=dev, dtype=torch.float64)
    return (torch.exp(u * torch.log(torch.tensor(float(V), device=dev, dtype=torch.float64))) - 1).long().clamp_(0, V - 1)


# ---- postings: unique (term, doc) pairs sorted by term then doc, tf = multiplicity ----
keys = []
for s in range(0, N, 1 << 19):
    e = min(N, s + (1 << 19))
    docs = torch.arange(s, e, device=dev).repeat_interleave(TOK)
    keys.append(zipf_terms(docs.numel()) * N + docs)
key, tf = torch.unique(torch.cat(keys), return_counts=True)
del keys
term, doc = key // N, (key % N).int()
del key
df = torch.bincount(term, minlength=V)
indptr = torch.zeros(V + 1, dtype=torch.int64, device=dev)
indptr[1:] = torch.cumsum(df, 0)
nnz = doc.numel()
tf32 = tf.float()
doc_len = torch.bincount(doc.long(), weights=tf.double(), minlength=N)
avdl = doc_len.mean()
k1, b = 1.2, 0.75
idf = torch.log(N / df.clamp(min=1).double())
doc_norm = k1 * (1 - b + b * doc_len / avdl)
val = torch.empty(nnz, dtype=torch.float64, device=dev)
L = _lib.lib()
_lib.check(L.ccr_bm25_build_impacts(indptr.data_ptr(), doc.data_ptr(), tf32.data_ptr(), idf.data_ptr(),
                                    doc_norm.data_ptr(), k1, V, nnz, val.data_ptr(), _stream_ptr(dev)))

# ---- queries: 3..10 distinct Zipf terms each ----
rows, q_indptr = [], [0]
lens = torch.randint(3, 11, (Q,), generator=torch.Generator().manual_seed(4)).tolist()
for n in lens:
    t = torch.unique(zipf_terms(n))
    rows.append(t)
    q_indptr.append(q_indptr[-1] + t.numel())
q_terms = torch.cat(rows).int()
q_indptr = torch.tensor(q_indptr, dtype=torch.int64, device=dev)
longest = max(r.numel() for r in rows)
postings_touched = int(df[q_terms.long()].sum())
out_s = torch.empty((Q, K), dtype=torch.float32, device=dev)
out_i = torch.empty((Q, K), dtype=torch.int64, device=dev)
need = L.ccr_bm25_topk_workspace_bytes(Q, N, K)
ws = torch.empty(need, dtype=torch.uint8, device=dev)

# ---- reference rows for the cross-check: torch, same term order, float64, ranked as float32 ----
CHECK_ROWS = list(range(4)) + [Q // 2, Q - 1]
want = []
for r in CHECK_ROWS:
    sc = torch.zeros(N, dtype=torch.float64, device=dev)
    for t in rows[r].tolist():
        lo, hi = int(indptr[t]), int(indptr[t + 1])
        sc[doc[lo:hi].long()] += val[lo:hi]
    ws_, wi_ = sc.float().sort(descending=True, stable=True)
    want.append((ws_[:K].clone(), wi_[:K].clone()))
del sc, ws_, wi_


def bench_one(head_frac, kern):
    """head_frac: terms present in >= that share of the docs get a dense float64 row (0 = postings only)."""
    head_slot = head_rows = None
    n_head = 0
    if head_frac > 0:
        head = torch.nonzero(df >= head_frac * N).flatten()
        head = head[torch.argsort(df[head], descending=True)][:64].int().contiguous()
        n_head = head.numel()
        if n_head:
            head_slot = torch.empty(V, dtype=torch.int32, device=dev)
            head_rows = torch.empty((n_head, L.ccr_bm25_head_row_pitch(N)), dtype=torch.float64, device=dev)
            _lib.check(L.ccr_bm25_build_head_rows(indptr.data_ptr(), doc.data_ptr(), val.data_ptr(), head.data_ptr(),
                                                  n_head, V, N, head_slot.data_ptr(), head_rows.data_ptr(),
                                                  _stream_ptr(dev)))
    hs_ptr = head_slot.data_ptr() if n_head else None
    hr_ptr = head_rows.data_ptr() if n_head else None
    is_head = (head_slot[q_terms.long()] >= 0) if n_head else torch.zeros_like(q_terms, dtype=torch.bool)
    head_visits = int(is_head.sum())                         # (query, head term) pairs
    bytes_moved = int(df[q_terms.long()][~is_head].sum()) * 12 + head_visits * N * 8

    def run():
        _lib.check(L.ccr_bm25_topk(indptr.data_ptr(), doc.data_ptr(), val.data_ptr(), hs_ptr, hr_ptr,
                                   q_indptr.data_ptr(), q_terms.data_ptr(), longest, Q, N, K, out_s.data_ptr(),
                                   out_i.data_ptr(), ws.data_ptr(), need, _stream_ptr(dev)))

    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 5
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    bad = 0
    for r, (ws_, wi_) in zip(CHECK_ROWS, want):
        bad += int((out_i[r] != wi_).sum()) + int((out_s[r] != ws_).sum())
    print(json.dumps({
        "what": "BM25 fused postings accumulation + top-k (ccr_bm25_topk), synthetic Zipf postings",
        "n_docs": N, "n_terms": V, "nnz": nnz, "queries": Q, "k": K, "ms_per_batch": ms,
        "queries_per_s": Q / ms * 1e3, "postings_touched_per_batch": postings_touched,
        "postings_GBps": postings_touched * 12 / ms / 1e6, "mismatches_vs_torch_f64": bad,
        "kernel": kern, "head_df_fraction": head_frac, "head_terms": n_head, "head_row_visits": head_visits,
        "bytes_moved_GBps": bytes_moved / ms / 1e6,
        "reference_cpu": "21 min 19 s for 3,452 NQ queries (al_demo_nq.ipynb:353) = 2.70 queries/s",
    }), flush=True)


def use_lib(path):
    """Builder A/B: switch to another build of libccr_b200 inside this process (same CUDA context)."""
    global L, need, ws
    _lib.LIB_PATH = path
    _lib._lib = None
    L = _lib.lib()
    need = L.ccr_bm25_topk_workspace_bytes(Q, N, K)   # builds may plan different candidate buffers
    ws = torch.empty(need, dtype=torch.uint8, device=dev)


# BM25_LIBS: comma-separated alternative builds of the library to time (default: the in-tree one)
# BM25_HEAD_FRAC: comma-separated list of head-term document-frequency fractions to measure (0 = postings only)
# BM25_KERNELS: comma-separated list of kernels: auto (warp-private for these short queries) | blockwide
for libpath in os.environ.get("BM25_LIBS", "").split(","):
    if libpath:
        use_lib(libpath)
    for kern in os.environ.get("BM25_KERNELS", "auto").split(","):
        os.environ.pop("CCR_BM25_BLOCKWIDE", None)
        if kern == "blockwide":
            os.environ["CCR_BM25_BLOCKWIDE"] = "1"
        _lib.reload_env()
        for frac in os.environ.get("BM25_HEAD_FRAC", "0.25").split(","):
            bench_one(float(frac), kern + (":" + os.path.basename(libpath) if libpath else ""))
