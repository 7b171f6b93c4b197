"""CPU oracle for CCR's candidate score-and-rank path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module -- and there only as the checker or as
the timed *reference* arm, never as the product path.  The product
(``ccr_b200``) never imports it and fails loudly when the CUDA library is missing.

What is restated here (reference paths relative to /root/reference):

* ``ranking_ref``              <- scripts/ms_marco_eval.py:189-235  (``ranking``)
* ``generate_embeddings_ref``  <- scripts/ms_marco_eval.py:123-152
* ``cos_sim_ref``              <- scripts/ms_marco_eval.py:155-162, src/ccrec/models/bbpr.py:485-492
* ``lazy_score_dense_ref``     <- src/rime_lite/util/score_array.py:173-174,226-227,291-293
                                  (MatMul(LazyDense, LazyDense.T) [+ LazySparse]).as_tensor("cpu")
* ``assign_topk_ref``          <- src/rime_lite/util/__init__.py:117-152 (``_assign_topk``)
* ``argsort_ref``              <- src/rime_lite/util/__init__.py:158-184 (``_argsort``)
* ``transform_scores_ref``     <- src/ccrec/models/bbpr.py:528-545 (tile loop of ``BertBPR.transform``)
* ``BM25Ref``                  <- scripts/bm_25.py:9-45 (``BM25.fit/cache/transform``)
* ``ranking_bm25_ref``         <- scripts/ms_marco_eval.py:165-186 (``ranking_bm25``)
* ``al0_requests_ref``         <- scripts/al_0_rank.py:136-218 (candidates, id_track, request_orig/perm)

The arithmetic of all of these lives in PyTorch (third-party, version unpinned by the
reference's setup.py:10-17): ``@``/``mm``, ``F.normalize``, ``Tensor.sort``,
``Tensor.topk``.  The restatements call the same torch ops on CPU tensors.

Pinning: the reference ships NO test, golden vector or fixture for this path
(test/ only holds test_dawid_skene.py).  The oracle is therefore pinned against
outputs of the reference's own code executed in the build container
(tests/golden/make_golden.py imports the unmodified reference -- or, for the
non-importable al_0_rank.py, executes its statements read from the file -- and stores
the outputs in tests/golden/*.npz; tests/test_oracle_golden.py and
tests/test_al_rank_cpu.py replay them).

``score_topk_ref`` (and ``score_topk_ref_device``, the same arithmetic with stock torch fp32 ops
on the GPU so that the full 8.84 M-row shapes finish in seconds) is the arbiter for the CUDA kernels: fp32 (or fp64 when an
additive float64 prior is present, as in the reference) scores from the
*bf16-rounded* inputs, chunked over the corpus with a running top-k so it scales,
ties broken by lowest id.  ``check_topk`` implements the tolerance rule of
BASELINE.json's north_star (scores within 1e-2 relative; id sets equal except
for swaps among items whose score lies within that tolerance of the k-th score).
"""
from __future__ import annotations

import math
import os

import numpy as np
import scipy.sparse as sps
import torch

MASK_NONE, MASK_SET, MASK_ADD = 0, 1, 2
RANKING_TOPN = 1001  # scripts/ms_marco_eval.py:230


# ----------------------------------------------------------------------------------------
# restatements of the reference functions
# ----------------------------------------------------------------------------------------
def cos_sim_ref(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """scripts/ms_marco_eval.py:155-162."""
    if a.dim() == 1:
        a = a.unsqueeze(0)
    if b.dim() == 1:
        b = b.unsqueeze(0)
    a_norm = torch.nn.functional.normalize(a, p=2, dim=1)
    b_norm = torch.nn.functional.normalize(b, p=2, dim=1)
    return torch.mm(a_norm, b_norm.transpose(0, 1))


def generate_embeddings_ref(data_indices, data_dic, embedding_func, batch_size):
    """scripts/ms_marco_eval.py:123-152 without the progress prints / torch.save."""
    num = len(data_indices)
    out = []
    with torch.no_grad():
        for step in range(math.ceil(num / batch_size)):
            indices = data_indices[step * batch_size : (step + 1) * batch_size]
            out.append(torch.as_tensor(embedding_func([data_dic[i] for i in indices])).to("cpu"))
    return torch.vstack(out)


def _autocast_fp16_mm(a, b_t):
    """What ``a @ b_t`` / ``torch.mm`` compute for CUDA tensors under ``torch.cuda.amp.autocast()``
    (scripts/al_0_rank.py:125 wraps the ranking call in it): operands cast to fp16, products
    accumulated in fp32 by the tensor cores, result rounded to fp16 -- restated with CPU fp32 ops.
    The accumulation order of cuBLAS is not reproduced, so a result may differ from the GPU's by one
    fp16 ulp when the fp32 sum falls next to a rounding boundary."""
    return (a.half().float() @ b_t.half().float()).half().float()


def ranking_ref(corpus, queries, embedding_func, batch_size, block_dict=None, sim_type=None,
                topn=RANKING_TOPN, stable=True, autocast_fp16=False):
    """scripts/ms_marco_eval.py:189-235 with its four CUDA calls removed.

    ``stable=True`` makes the per-row sort stable (ties -> lowest corpus position first),
    which is one of the orders the reference's unstable ``sort`` may legally produce.
    ``autocast_fp16=True`` restates the reference's REAL GPU arithmetic (fp16 autocast matmul,
    fp32 normalisation, scores stored back into the fp32 host matrix); pinned by
    tests/golden/ranking_cuda_autocast_*.npz, produced by the unmodified function on a B200.
    """
    import pandas as pd

    if sim_type is None:
        sim_type = os.environ.get("CCREC_SIM_TYPE", "cos")
    num_queries, num_passages = len(queries), len(corpus)
    queries_ids, corpus_ids = list(queries.keys()), list(corpus.keys())
    queries_embeddings = generate_embeddings_ref(queries_ids, queries, embedding_func, batch_size)
    passage_embeddings = generate_embeddings_ref(corpus_ids, corpus, embedding_func, batch_size)

    ranking_profile = {}
    ranking_matrix = torch.zeros(num_queries, num_passages)          # :204
    for step in range(math.ceil(num_passages / batch_size)):         # :206-218
        pb = passage_embeddings[step * batch_size : (step + 1) * batch_size]
        if autocast_fp16:
            if sim_type == "cos":   # F.normalize stays fp32 under autocast, torch.mm runs in fp16
                scores = _autocast_fp16_mm(normalize_rows_ref(queries_embeddings), normalize_rows_ref(pb).T)
            else:
                scores = _autocast_fp16_mm(queries_embeddings, pb.T)
        elif sim_type == "cos":
            scores = cos_sim_ref(queries_embeddings, pb)
        else:
            scores = queries_embeddings @ pb.T
        ranking_matrix[:, step * batch_size : step * batch_size + pb.shape[0]] = scores
    corpus_index = pd.Index(corpus_ids)  # hoisted out of the loop; same get_indexer result
    for step, qid in enumerate(queries_ids):                         # :221-234
        scores = ranking_matrix[step]
        if block_dict is not None:
            block_ind = corpus_index.get_indexer(block_dict[qid])
            assert -1 not in block_ind, "block id not found"
            scores[block_ind] = -1e6                                 # :227 assignment
        ordered_scores, ordering = scores.sort(descending=True, stable=stable)
        ordered_scores, ordering = ordered_scores[0:topn], ordering[0:topn]
        ordered_pids = [corpus_ids[idx] for idx in ordering]
        ranking_profile[qid] = dict(zip(ordered_pids, ordered_scores.numpy().tolist()))
    return ranking_profile


def ranking_core_ref(queries_embeddings, passage_embeddings, batch_size=512, block_rows=None, sim_type="dot",
                     topn=RANKING_TOPN):
    """The arithmetic core of ``ranking`` (scripts/ms_marco_eval.py:203-230) on tensors, CPU only:
    tile matmul into the host Q x N matrix, ``scores[block_ind] = -1e6``, full per-row sort,
    first ``topn``.  ``block_rows``: per query, an int64 array of blocked corpus positions.
    Returns (scores [Q, topn] f32, positions [Q, topn] i64).  This is what ``bench.py`` times as
    the reference's CPU path."""
    Q, N = queries_embeddings.shape[0], passage_embeddings.shape[0]
    ranking_matrix = torch.zeros(Q, N)
    for step in range(math.ceil(N / batch_size)):
        pb = passage_embeddings[step * batch_size : (step + 1) * batch_size]
        if sim_type == "cos":
            scores = cos_sim_ref(queries_embeddings, pb)
        else:
            scores = queries_embeddings @ pb.T
        ranking_matrix[:, step * batch_size : step * batch_size + pb.shape[0]] = scores
    kk = min(topn, N)
    out_s = torch.empty(Q, kk)
    out_i = torch.empty(Q, kk, dtype=torch.int64)
    for step in range(Q):
        scores = ranking_matrix[step]
        if block_rows is not None:
            scores[torch.as_tensor(block_rows[step], dtype=torch.int64)] = -1e6
        ordered_scores, ordering = scores.sort(descending=True)
        out_s[step], out_i[step] = ordered_scores[:kk], ordering[:kk]
    return out_s, out_i


def lazy_score_dense_ref(U, V, prior=None):
    """What ``(LazyDense(U) @ LazyDense(V).T [+ prior_csr]).as_tensor("cpu")`` evaluates to.

    score_array.py:226-227 (``torch.as_tensor(self.c)``), :291-293 (``op(*children)``),
    :173-174 (sparse -> dense).  fp32 @ fp32 -> fp32; adding a float64 CSR promotes to float64.
    """
    s = torch.as_tensor(np.asarray(U)) @ torch.as_tensor(np.asarray(V)).T
    if prior is not None:
        s = s + torch.as_tensor(sps.csr_matrix(prior).toarray())
    return s


def assign_topk_ref(S, k, tie_breaker=0.0, seed=None):
    """src/rime_lite/util/__init__.py:117-152 on an already-dense matrix ``S``.

    With ``tie_breaker=0`` this is ``S.topk(k).indices`` -> CSR of ones with the indices
    in top-k order (:145-152).  Raises like torch when k > n_cols.
    """
    s = torch.as_tensor(np.asarray(S)) if not torch.is_tensor(S) else S
    if tie_breaker:
        g = torch.Generator().manual_seed(0 if seed is None else seed)
        s = s + torch.rand(*s.shape, generator=g) * tie_breaker
    indices = s.topk(k).indices.cpu().numpy()
    return sps.csr_matrix(
        (np.ones(indices.size), np.ravel(indices), np.arange(0, indices.size + 1, indices.shape[1])),
        shape=tuple(s.shape),
    )


def argsort_ref(S):
    """src/rime_lite/util/__init__.py:158-184 with tie_breaker=0: flat descending argsort."""
    S = np.asarray(S)
    ind = np.argsort(-S.reshape(-1), kind="stable")
    return np.unravel_index(ind, S.shape)


def transform_scores_ref(all_emb, i_to_ptr, j_to_ptr, batch_size, sim_type):
    """src/ccrec/models/bbpr.py:528-545: users x items dense fp32 score matrix by item tiles."""
    all_emb = torch.as_tensor(all_emb)
    user_embedding = all_emb[i_to_ptr]
    out = torch.zeros(len(i_to_ptr), len(j_to_ptr))
    for step in range(int(np.ceil(len(j_to_ptr) / batch_size))):
        item_ids = j_to_ptr[step * batch_size : (step + 1) * batch_size]
        ib = all_emb[item_ids]
        scores = cos_sim_ref(user_embedding, ib) if sim_type == "cos" else user_embedding @ ib.T
        out[:, step * batch_size : step * batch_size + ib.shape[0]] = scores
    return out


class BM25Ref:
    """scripts/bm_25.py:9-45 restated on the raw CSC arrays.

    The vocabulary / counting is sklearn's (``TfidfVectorizer(norm=None, smooth_idf=False)``
    fitted on the corpus, counts from its ``CountVectorizer`` base, bm_25.py:11,23,37); the BM25
    arithmetic (bm_25.py:39-45) is written out per query term instead of through scipy's
    sparse/dense operators.  Those operators live in scipy (third party, unpinned by the
    reference; 1.18.1 in this image, whose evaluation order the goldens capture):
    ``numer / denom`` with a sparse numerator and a dense denominator is evaluated as
    ``numer.multiply(1 / denom)`` (scipy/sparse/_base.py ``_divide``) and stays sparse, and
    ``.sum(1)`` of that sparse matrix adds each row's stored entries in column order.  So for a
    query with distinct vocabulary terms t_1 < ... < t_T

        score[d] = sum_j  ((tf_jd * idf_j) * (k1 + 1)) * (1 / (tf_jd + k1 * (1 - b + b * len_d / avdl)))

    over the terms present in d, added left to right in float64.
    """

    def __init__(self, b=0.75, k1=1.6):
        from sklearn.feature_extraction.text import TfidfVectorizer

        self.vectorizer = TfidfVectorizer(norm=None, smooth_idf=False)
        self.b, self.k1 = b, k1

    def counts(self, texts):
        from sklearn.feature_extraction.text import CountVectorizer

        return CountVectorizer.transform(self.vectorizer, texts)

    def fit(self, X):
        self.vectorizer.fit(X)
        self.cache(X)
        self.avdl = self.doc_len.mean()
        return self

    def cache(self, X):
        self.csc = self.counts(X).tocsc()
        self.csc.sort_indices()
        self.doc_len = np.asarray(self.csc.sum(1)).ravel()
        return self

    def query_terms(self, q):
        row = self.counts([q])
        return row.indices  # distinct vocabulary ids, ascending (sklearn sorts CSR indices)

    def transform(self, q):
        terms = self.query_terms(q)
        n = self.csc.shape[0]
        norm = self.k1 * (1 - self.b + self.b * self.doc_len / self.avdl)
        idf = self.vectorizer._tfidf.idf_ - 1.0
        scores = np.zeros(n, dtype=np.float64)
        for t in terms:
            lo, hi = self.csc.indptr[t], self.csc.indptr[t + 1]
            docs = self.csc.indices[lo:hi]
            tf = self.csc.data[lo:hi].astype(np.float64)
            scores[docs] += ((tf * idf[t]) * (self.k1 + 1)) * (1.0 / (tf + norm[docs]))
        return scores


def ranking_bm25_ref(corpus, queries, topn=RANKING_TOPN, stable=True):
    """scripts/ms_marco_eval.py:165-186: fit on the corpus texts, per query score all docs, cast
    to float32, sort descending, keep the first 1001.  ``stable`` orders exact ties by corpus
    position (the reference's sort is unstable; ties are unspecified)."""
    model = BM25Ref(b=0.75, k1=1.2).fit(list(corpus.values()))
    corpus_ids = list(corpus.keys())
    profile = {}
    for qid, text in queries.items():
        solution = torch.Tensor(model.transform(text))
        scores, ordering = solution.sort(descending=True, stable=stable)
        scores, ordering = scores[0:topn], ordering[0:topn]
        profile[qid] = dict(zip([corpus_ids[i] for i in ordering], scores.numpy().tolist()))
    return profile


def al0_requests_ref(ranking_profile, ranking_profile_bm25, corpus, queries, qids_split, step,
                     number_of_qid_split_batch, n_repeats, repeat_seed, landing_image=None, display_length=250):
    """scripts/al_0_rank.py:136-218 as a function of its inputs.  Returns (header, rows,
    permuted_rows, id_track); the CSV files are ``pd.DataFrame(rows, columns=header).to_csv(index=False)``.

    :138 ranks_rng = RandomState(STEP); :165-167 only queries of split STEP % n; :169-177 dense
    top-2 then BM25 passages until three; :179-182 random corpus passages (one ``choice`` per
    attempt) until four; :184-194 row + id_track; :204-215 N_REPEATS passes, one ``permutation(4)``
    per row applied to passages, pids and images alike."""
    import re

    ranks_rng = np.random.RandomState(step)
    corpus_keys = list(corpus.keys())
    header = ["query"] + [f"passage-{i}" for i in range(1, 5)] + ["qid"] + [f"pid-{i}" for i in range(1, 5)]
    if landing_image is not None:
        header += ["img-q"] + [f"img-{i}" for i in range(1, 5)]
    split = qids_split[step % number_of_qid_split_batch]
    rows, id_track = [], {}
    for qid in ranking_profile:
        if qid not in split:
            continue
        cands = list(ranking_profile[qid].keys())[0:2]
        for pid in ranking_profile_bm25[qid].keys():
            if len(cands) == 3:
                break
            if pid not in cands:
                cands.append(pid)
        while len(cands) < 4:
            pid = corpus_keys[ranks_rng.choice(len(corpus_keys))]
            if pid not in cands:
                cands.append(pid)
        passages = [re.sub(r"[^a-zA-Z0-9 ,:.;?$!()&\[\]]", "", corpus[pid])[:display_length] for pid in cands]
        row = [queries[qid]] + passages + [f"q_{qid}"] + [f"p_{c}" for c in cands]
        if landing_image is not None:
            row = row + [landing_image[qid]] + [landing_image[c] for c in cands]
        rows.append(row)
        id_track[queries[qid]] = f"q_{qid}"
        for pid, passage in zip(cands, passages):
            id_track[passage] = f"p_{pid}"
    rng = np.random.RandomState(repeat_seed)
    permuted = []
    for _ in range(n_repeats):
        for row in rows:
            ind = rng.permutation(4)
            out = [row[0]] + [row[1 + i] for i in ind] + [row[5]] + [row[6 + i] for i in ind]
            if len(row) > 10:
                out = out + [row[10]] + [row[11 + i] for i in ind]
            permuted.append(out)
    return header, rows, permuted, id_track


# ----------------------------------------------------------------------------------------
# the kernel arbiter
# ----------------------------------------------------------------------------------------
def bf16_round(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


def normalize_rows_ref(x: torch.Tensor) -> torch.Tensor:
    """F.normalize(p=2, dim=1, eps=1e-12) in fp32 (ms_marco_eval.py:160-161)."""
    return torch.nn.functional.normalize(x.float(), p=2, dim=1)


def canonical_mask(indptr, cols, vals, n_cols, mode):
    """Sort columns per row and merge duplicates (sum for add, last-wins==same value for set)."""
    indptr = np.asarray(indptr, dtype=np.int64)
    m = sps.csr_matrix((np.asarray(vals, dtype=np.float64), np.asarray(cols, dtype=np.int64), indptr),
                       shape=(len(indptr) - 1, n_cols))
    if mode == MASK_SET:
        # duplicates under assignment: keep one entry (all carry the same value in the reference)
        coo = m.tocoo()
        key = coo.row.astype(np.int64) * n_cols + coo.col
        _, first = np.unique(key, return_index=True)
        m = sps.csr_matrix((np.asarray(vals, dtype=np.float64)[first], (coo.row[first], coo.col[first])),
                           shape=m.shape)
    else:
        m.sum_duplicates()
    m.sort_indices()
    return m.indptr.astype(np.int64), m.indices.astype(np.int32), m.data.astype(np.float64)


def score_topk_ref(Q, P, k, mask=None, mode=MASK_NONE, sim="dot", round_bf16=True, chunk=1 << 18,
                   id_offset=0, return_f64=False):
    """Exact top-k of ``Q @ P.T`` (+ sparse set/add mask) per row, ties -> lowest id.

    Q [B,D], P [N,D] float tensors.  ``round_bf16`` rounds both to bf16 first (after the
    fp32 L2-normalisation when ``sim == 'cos'``), matching what the device table stores.
    ``mask`` = (indptr[B+1], cols[nnz], vals[nnz]) over the N columns.  ``mode``:
    MASK_SET assigns ``vals`` (ranking(): -1e6, ms_marco_eval.py:227); MASK_ADD adds them in
    float64 (rime_lite prior_score, dataset/base.py:234,279-282).
    Returns (scores float32|float64 [B,k] descending, ids int64 [B,k] + id_offset).
    """
    Q = torch.as_tensor(Q).float()
    P = torch.as_tensor(P).float()
    B, N = Q.shape[0], P.shape[0]
    if k > N:
        raise RuntimeError("selected index k out of range")
    if sim == "cos":
        Q, P = normalize_rows_ref(Q), normalize_rows_ref(P)
    if round_bf16:
        Q, P = bf16_round(Q), bf16_round(P)
    use64 = mode == MASK_ADD
    dt = torch.float64 if use64 else torch.float32
    best_s = torch.full((B, 0), 0, dtype=dt)
    best_i = torch.zeros((B, 0), dtype=torch.int64)
    if mask is not None and mode != MASK_NONE:
        indptr, cols, vals = canonical_mask(mask[0], mask[1], mask[2], N, mode)
        mcsr = sps.csr_matrix((vals, cols, indptr), shape=(B, N))
    else:
        mcsr = None
    old = torch.backends.cuda.matmul.allow_tf32
    for s0 in range(0, N, chunk):
        s1 = min(N, s0 + chunk)
        sc = (Q @ P[s0:s1].T).to(dt)
        if mcsr is not None:
            sub = mcsr[:, s0:s1].tocoo()
            r = torch.as_tensor(sub.row, dtype=torch.int64)
            c = torch.as_tensor(sub.col, dtype=torch.int64)
            v = torch.as_tensor(sub.data, dtype=dt)
            if mode == MASK_SET:
                sc[r, c] = v
            else:
                sc[r, c] = sc[r, c] + v
        ids = torch.arange(s0, s1, dtype=torch.int64).expand(B, -1)
        cat_s = torch.cat([best_s, sc], dim=1)
        cat_i = torch.cat([best_i, ids], dim=1)
        # stable descending sort == ties -> earlier position == lower id (best_* precede and
        # hold lower ids than the current chunk)
        kk = min(k, cat_s.shape[1])
        o = torch.sort(cat_s, dim=1, descending=True, stable=True)
        best_s = o.values[:, :kk].contiguous()
        best_i = torch.gather(cat_i, 1, o.indices[:, :kk]).contiguous()
    torch.backends.cuda.matmul.allow_tf32 = old
    out_s = best_s if (return_f64 or not use64) else best_s
    if not return_f64:
        out_s = out_s.to(torch.float32)
    return out_s, best_i + id_offset


def score_topk_ref_device(Qe, Pe, k, mask=None, mode=MASK_NONE, chunk=1 << 17, id_offset=0, n_items=None):
    """``score_topk_ref`` at full corpus sizes: the same arithmetic (fp32 products of the
    bf16-rounded inputs, float64 where an additive prior is present, running top-k, ties -> lowest
    id) evaluated chunk by chunk with stock torch ops on the device the ENCODED operands live on.

    ``Qe`` [B, >=D], ``Pe`` [N, >=D]: the encoded (bf16-representable) queries / table rows, any
    float dtype; they are widened to fp32 per chunk and multiplied with TF32 disabled, so this is an
    fp32 reference of the kernels' bf16 x bf16 -> fp32 contraction, independent of them.
    Returns (scores float32 | float64 for MASK_ADD, ids int64) on the CPU.
    """
    dev = Qe.device
    B = Qe.shape[0]
    N = Pe.shape[0] if n_items is None else int(n_items)
    if k > N:
        raise RuntimeError("selected index k out of range")
    use64 = mode == MASK_ADD
    dt = torch.float64 if use64 else torch.float32
    old_tf32 = torch.backends.cuda.matmul.allow_tf32
    old_prec = torch.get_float32_matmul_precision()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")
    try:
        Qf = Qe.float()
        m_r = m_c = m_v = None
        if mask is not None and mode != MASK_NONE:
            indptr, cols, vals = canonical_mask(mask[0], mask[1], mask[2], N, mode)
            m_r = torch.as_tensor(np.repeat(np.arange(B, dtype=np.int64), np.diff(indptr)), device=dev)
            m_c = torch.as_tensor(cols.astype(np.int64), device=dev)
            m_v = torch.as_tensor(vals, dtype=dt, device=dev)
        best_s = torch.empty((B, 0), dtype=dt, device=dev)
        best_i = torch.empty((B, 0), dtype=torch.int64, device=dev)
        for s0 in range(0, N, chunk):
            s1 = min(N, s0 + chunk)
            sc = (Qf @ Pe[s0:s1].float().T).to(dt)
            if m_c is not None:
                sel = (m_c >= s0) & (m_c < s1)
                r, c, v = m_r[sel], m_c[sel] - s0, m_v[sel]
                if mode == MASK_SET:
                    sc[r, c] = v
                else:
                    sc[r, c] = sc[r, c] + v
            kk = min(k, s1 - s0)
            tv, ti = torch.topk(sc, kk, dim=1, sorted=True)
            # torch.topk does not promise the LOWEST ids among values tied with its kk-th: rows where
            # the chunk holds more copies of that value than were selected are redone by a stable sort
            vk = tv[:, -1:]
            redo = torch.nonzero((sc == vk).sum(1) != (tv == vk).sum(1)).flatten()
            if redo.numel():
                o = torch.sort(sc[redo], dim=1, descending=True, stable=True)
                tv[redo], ti[redo] = o.values[:, :kk], o.indices[:, :kk]
            cat_s = torch.cat([best_s, tv], dim=1)
            cat_i = torch.cat([best_i, ti + s0], dim=1)
            # order by (score descending, id ascending): sort by id, then a stable sort by score
            o1 = torch.sort(cat_i, dim=1, stable=True)
            cat_s = torch.gather(cat_s, 1, o1.indices)
            o2 = torch.sort(cat_s, dim=1, descending=True, stable=True)
            keep = min(k, cat_s.shape[1])
            best_s = o2.values[:, :keep].contiguous()
            best_i = torch.gather(o1.values, 1, o2.indices[:, :keep]).contiguous()
            del sc, tv, ti
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old_tf32
        torch.set_float32_matmul_precision(old_prec)
    return best_s.cpu(), (best_i + id_offset).cpu()


def full_scores_ref(Q, P, mask=None, mode=MASK_NONE, sim="dot", round_bf16=True):
    """Dense [B,N] score matrix under the same conventions as ``score_topk_ref`` (small cases)."""
    Q = torch.as_tensor(Q).float()
    P = torch.as_tensor(P).float()
    if sim == "cos":
        Q, P = normalize_rows_ref(Q), normalize_rows_ref(P)
    if round_bf16:
        Q, P = bf16_round(Q), bf16_round(P)
    s = Q @ P.T
    if mask is not None and mode != MASK_NONE:
        indptr, cols, vals = canonical_mask(mask[0], mask[1], mask[2], P.shape[0], mode)
        m = sps.csr_matrix((vals, cols, indptr), shape=s.shape).tocoo()
        r, c = torch.as_tensor(m.row, dtype=torch.int64), torch.as_tensor(m.col, dtype=torch.int64)
        if mode == MASK_SET:
            s[r, c] = torch.as_tensor(m.data, dtype=torch.float32)
        else:
            s = s.double()
            s[r, c] += torch.as_tensor(m.data, dtype=torch.float64)
    return s


def check_topk(got_scores, got_ids, full_scores=None, ref_scores=None, ref_ids=None, rtol=1e-2,
               atol=1e-6, ordered_slack=True):
    """The north_star tolerance rule.  Returns a list of human-readable violations (empty == ok).

    Either ``full_scores`` [B,N] (small cases: every returned id is checked against its true
    score) or ``ref_scores``/``ref_ids`` [B,k] from ``score_topk_ref`` must be given.

    * scores: |s - s_ref| <= rtol*|s_ref| + atol for the score of every returned id
      (vs the full matrix) or rank-wise (vs the reference list);
    * ids: distinct; set-equal to the reference except for items whose reference score lies
      within the tolerance of the reference k-th score;
    * order: returned scores non-increasing.
    """
    errs = []
    gs = np.asarray(got_scores, dtype=np.float64)
    gi = np.asarray(got_ids, dtype=np.int64)
    B, k = gi.shape
    for b in range(B):
        if len(set(gi[b].tolist())) != k:
            errs.append(f"row {b}: duplicate ids")
            continue
        if np.any(np.diff(gs[b]) > rtol * np.abs(gs[b][1:]) + atol):
            errs.append(f"row {b}: scores not descending")
        if full_scores is not None:
            row = np.asarray(full_scores[b], dtype=np.float64)
            true = row[gi[b]]
            bad = np.abs(gs[b] - true) > rtol * np.abs(true) + atol
            if bad.any():
                j = int(np.argmax(bad))
                errs.append(f"row {b}: score of id {gi[b, j]} is {gs[b, j]} vs {true[j]}")
            kth = np.sort(row)[::-1][k - 1]
            tol = rtol * abs(kth) + atol
            # every returned id must be >= kth - tol ; every item > kth + tol must be returned
            if (true < kth - tol).any():
                j = int(np.argmin(true))
                errs.append(f"row {b}: id {gi[b, j]} (score {true[j]}) below k-th {kth}")
            must = np.nonzero(row > kth + tol)[0]
            missing = np.setdiff1d(must, gi[b])
            if missing.size:
                errs.append(f"row {b}: missing ids {missing[:5].tolist()} above k-th {kth}")
        else:
            rs = np.asarray(ref_scores[b], dtype=np.float64)
            ri = np.asarray(ref_ids[b], dtype=np.int64)
            bad = np.abs(gs[b] - rs) > rtol * np.abs(rs) + atol
            if bad.any():
                j = int(np.argmax(bad))
                errs.append(f"row {b}: rank {j} score {gs[b, j]} vs ref {rs[j]}")
            kth = rs[-1]
            tol = rtol * abs(kth) + atol
            sure = ri[rs > kth + tol]
            missing = np.setdiff1d(sure, gi[b])
            if missing.size:
                errs.append(f"row {b}: missing ids {missing[:5].tolist()}")
            extra = np.setdiff1d(gi[b], ri)
            if extra.size:
                # extras are only allowed as near-tie swaps: their returned score must be near kth
                pos = np.nonzero(np.isin(gi[b], extra))[0]
                if (gs[b][pos] < kth - tol).any():
                    errs.append(f"row {b}: extra ids {extra[:5].tolist()} below k-th {kth}")
        if len(errs) > 20:
            break
    return errs
