"""Drop-in for the reference's dense retrieval entry point.

``ranking(corpus, queries, embedding_func, batch_size, block_dict=None)`` keeps the signature,
inputs and return value of scripts/ms_marco_eval.py:189-235 -- ``{qid: {pid: score}}`` holding
the best 1001 (or N) passages per query in descending score order, block-listed passages
carrying -1e6 -- but never builds the Q x N host matrix nor sorts full rows: embeddings go
straight into a device-resident bf16 table and one fused kernel does score + mask + top-k.

Deliberate deviations (documented in DESIGN.md): scores come from bf16-rounded embeddings with
fp32 accumulation (the reference's GPU path under autocast is fp16-in/fp16-out,
scripts/al_0_rank.py:125); exact ties are ordered by corpus position (reference: unspecified).
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch

from . import engine
from .table import EmbeddingTable

RANKING_TOPN = 1001  # scripts/ms_marco_eval.py:230
BLOCK_VALUE = -1e6   # scripts/ms_marco_eval.py:227
QUERY_CHUNK = 8192   # query rows per fused call (bounds workspace and the [B,k] result)


def cos_sim(a: torch.Tensor, b: torch.Tensor):
    """scripts/ms_marco_eval.py:155-162 on the device path: fp32 normalise -> bf16 -> dense scores."""
    a = torch.as_tensor(a)
    b = torch.as_tensor(b)
    if a.dim() == 1:
        a = a.unsqueeze(0)
    if b.dim() == 1:
        b = b.unsqueeze(0)
    t = EmbeddingTable.from_tensor(b, normalize=True)
    return t.dense_scores(a)


def generate_embeddings(data_indices, data_dic, embedding_func, batch_size, embedding_size=768, name=None):
    """Reference-compatible signature (scripts/ms_marco_eval.py:123-152): returns a CPU fp32
    tensor.  Kept for callers that want the host copy; ``ranking`` uses
    ``generate_embeddings_device`` instead."""
    out = []
    with torch.no_grad():
        for step in range(math.ceil(len(data_indices) / batch_size)):
            idx = data_indices[step * batch_size : (step + 1) * batch_size]
            out.append(torch.as_tensor(embedding_func([data_dic[i] for i in idx])).to("cpu"))
    emb = torch.vstack(out) if out else torch.zeros(0, embedding_size)
    if name is not None:
        torch.save(emb, name)
    return emb


def generate_embeddings_device(data_indices, data_dic, embedding_func, batch_size, normalize=False,
                               device="cuda", id_offset=0, name=None):
    """Encoder batches -> device bf16 table without the host round trip."""
    table = None
    with torch.no_grad():
        for step in range(math.ceil(len(data_indices) / batch_size)):
            idx = data_indices[step * batch_size : (step + 1) * batch_size]
            emb = torch.as_tensor(embedding_func([data_dic[i] for i in idx]))
            if table is None:
                table = EmbeddingTable(len(data_indices), emb.shape[1], device=device, normalize=normalize,
                                       id_offset=id_offset)
            table.append(emb)
    if table is None:
        table = EmbeddingTable(0, 768, device=device, normalize=normalize, id_offset=id_offset)
    if name is not None:
        table.save(name)
    return table


def build_block_mask(queries_ids, corpus_ids, block_dict, device):
    """block_dict {qid: [pid, ...]} -> SparseMask(set -1e6).  Mirrors ms_marco_eval.py:225-227
    including the "block id not found" assertion, with the pid->position index built once and ONE
    lookup over all block lists instead of one index build + lookup per query."""
    import pandas as pd

    from itertools import chain

    index = pd.Index(corpus_ids)
    blocks = [block_dict[qid] for qid in queries_ids]  # KeyError for a query without an entry, like :225
    lengths = [len(b) for b in blocks]
    ind = index.get_indexer(list(chain.from_iterable(blocks))) if sum(lengths) else np.zeros(0, dtype=np.int64)
    assert -1 not in ind, "block id not found"
    return engine.SparseMask.from_flat(lengths, ind, len(corpus_ids), BLOCK_VALUE, engine.MASK_SET, device)


def ranking_tensors(query_table, passage_table, k, mask=None, algo=0):
    """Core of ``ranking``: resident tables -> (scores [Q,k] f32, positions [Q,k] i64) on the host."""
    Q = len(query_table)
    scores = np.empty((Q, k), dtype=np.float32)
    order = np.empty((Q, k), dtype=np.int64)
    for s in range(0, Q, QUERY_CHUNK):
        e = min(Q, s + QUERY_CHUNK)
        m = mask.rows(s, e) if mask is not None else None
        sc, ids = passage_table.search(query_table.data[s:e], k, mask=m, algo=algo, encoded=True)
        scores[s:e] = sc.cpu().numpy()
        order[s:e] = ids.cpu().numpy()
    return scores, order


def ranking(corpus, queries, embedding_func, batch_size, block_dict=None, device="cuda", algo=0):
    sim = os.environ["CCREC_SIM_TYPE"]  # KeyError when unset, like ms_marco_eval.py:212
    normalize = sim == "cos"
    queries_ids, corpus_ids = list(queries.keys()), list(corpus.keys())
    q_table = generate_embeddings_device(queries_ids, queries, embedding_func, batch_size, normalize, device)
    p_table = generate_embeddings_device(corpus_ids, corpus, embedding_func, batch_size, normalize, device)
    if len(queries_ids) == 0:
        return {}
    mask = None
    if block_dict is not None:
        print("using block_dict")
        mask = build_block_mask(queries_ids, corpus_ids, block_dict, p_table.device)
    k = min(RANKING_TOPN, len(corpus_ids))
    scores, order = ranking_tensors(q_table, p_table, k, mask, algo=algo)
    corpus_arr = np.asarray(corpus_ids, dtype=object)
    ranking_profile = {}
    for step, qid in enumerate(queries_ids):
        ranking_profile[qid] = dict(zip(corpus_arr[order[step]].tolist(), scores[step].tolist()))
    return ranking_profile


def mrr_at_k(order, corpus_ids, queries_ids, qrels, k_values=(1, 5, 10, 100)):
    """MRR@k of a ranking produced by ``ranking_tensors`` (positions [Q, k], descending), the
    metric scripts/al_0_rank.py:130-133 obtains from BEIR's ``evaluate_custom(..., metric="mrr")``:
    per query the reciprocal rank of the first relevant passage (qrels score > 0) within the top
    k, averaged over queries, rounded to 5 digits.  Vectorised over queries."""
    pos = {pid: i for i, pid in enumerate(corpus_ids)}
    Q, K = order.shape
    first = np.full(Q, np.inf)
    for qi, qid in enumerate(queries_ids):
        rel = [pos[p] for p, s in qrels.get(qid, {}).items() if s > 0 and p in pos]
        if rel:
            hit = np.nonzero(np.isin(order[qi], rel))[0]
            if hit.size:
                first[qi] = hit[0] + 1
    return {f"MRR@{k}": round(float(np.mean(np.where(first <= k, 1.0 / first, 0.0))), 5) for k in k_values}


def ranking_sharded(corpus, queries, embedding_func, batch_size, block_dict=None, group=None, device=None,
                    index_cls=None):
    """``ranking`` with the passage table row-sharded over the ranks of a torch.distributed group
    (one process per GPU): every rank encodes the queries and ONLY its own contiguous slice of the
    corpus (``shard_bounds``) straight into its device shard -- the encoder pass, which dominates the
    reference's wall clock, is split G ways and no rank ever holds the whole table -- then the
    fused local top-k, one all-gather and the G-way merge produce the global result.  Every rank
    returns the same ``{qid: {pid: score}}`` as the single-table call."""
    from .dist import ShardedIndex

    sim = os.environ["CCREC_SIM_TYPE"]  # KeyError when unset, like ms_marco_eval.py:212
    normalize = sim == "cos"
    queries_ids, corpus_ids = list(queries.keys()), list(corpus.keys())
    if len(queries_ids) == 0:
        return {}
    with torch.no_grad():
        q_emb = torch.cat([torch.as_tensor(embedding_func([queries[i] for i in queries_ids[s:s + batch_size]]))
                           for s in range(0, len(queries_ids), batch_size)])
        index = (index_cls or ShardedIndex)(len(corpus_ids), q_emb.shape[1], normalize=normalize, device=device,
                                            group=group)
        for s in range(index.lo, index.hi, batch_size):
            e = min(index.hi, s + batch_size)
            index.add_local(torch.as_tensor(embedding_func([corpus[i] for i in corpus_ids[s:e]])))
    mask = None
    if block_dict is not None:
        print("using block_dict")
        mask = build_block_mask(queries_ids, corpus_ids, block_dict, index.device)
    k = min(RANKING_TOPN, len(corpus_ids))
    scores = np.empty((len(queries_ids), k), dtype=np.float32)
    order = np.empty((len(queries_ids), k), dtype=np.int64)
    for s in range(0, len(queries_ids), QUERY_CHUNK):
        e = min(len(queries_ids), s + QUERY_CHUNK)
        sc, ids, _ = index.search(q_emb[s:e], k, mask=mask.rows(s, e) if mask is not None else None)
        scores[s:e] = sc.cpu().numpy()
        order[s:e] = ids.cpu().numpy()
    corpus_arr = np.asarray(corpus_ids, dtype=object)
    return {qid: dict(zip(corpus_arr[order[step]].tolist(), scores[step].tolist()))
            for step, qid in enumerate(queries_ids)}
