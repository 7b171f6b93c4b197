"""Seeded synthetic inputs shared by make_golden.py (which runs the real reference on
them) and by the tests (which replay them through the oracle and the CUDA path).

``np.random.RandomState`` (legacy MT19937 stream) is frozen by numpy's compatibility
policy, so regenerating from the seed reproduces the exact inputs the goldens were made
from; the goldens only store the reference's OUTPUTS.
"""
import numpy as np
import scipy.sparse as sps


def embeddings(seed, n, d, clustered=False):
    rs = np.random.RandomState(seed)
    if not clustered:
        return rs.standard_normal((n, d)).astype(np.float32)
    centres = rs.standard_normal((max(4, n // 16), d)).astype(np.float32)
    x = centres[rs.randint(0, len(centres), size=n)] + 0.5 * rs.standard_normal((n, d)).astype(np.float32)
    return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)


class TextTable:
    """A synthetic ``embedding_func``: text key -> fixed embedding row."""

    def __init__(self, table):
        import torch

        self.table = torch.as_tensor(table)
        self.calls = 0

    def __call__(self, texts):
        import torch

        self.calls += 1
        idx = torch.as_tensor([int(t.split("#")[1]) for t in texts], dtype=torch.int64)
        return self.table[idx]


def ranking_case(name):
    """-> dict(corpus, queries, table, batch_size, block_dict, sim_type)."""
    spec = RANKING_CASES[name] if name in RANKING_CASES else CUDA_RANKING_CASES[name]
    n, q, d = spec["n"], spec["q"], spec["d"]
    same = spec.get("queries_are_corpus", False)
    emb = embeddings(spec["seed"], n if same else n + q, d, spec.get("clustered", False))
    corpus = {f"p{i}": f"text#{i}" for i in range(n)}
    if same:
        queries = {f"p{i}": f"text#{i}" for i in range(q)}
    else:
        queries = {f"q{i}": f"text#{n + i}" for i in range(q)}
    block_dict = None
    if spec.get("brands"):
        rs = np.random.RandomState(spec["seed"] + 1)
        brand = rs.zipf(1.1, size=n) % spec["brands"]
        groups = {}
        for i, b in enumerate(brand):
            groups.setdefault(int(b), []).append(f"p{i}")
        if same:
            block_dict = {f"p{i}": groups[int(brand[i])] for i in range(q)}
        else:
            block_dict = {f"q{i}": groups[int(brand[rs.randint(n)])] for i in range(q)}
    return dict(corpus=corpus, queries=queries, table=emb, batch_size=spec["batch_size"],
                block_dict=block_dict, sim_type=spec["sim"])


RANKING_CASES = {
    # name: N passages, Q queries, dim, similarity, encoder/tile batch, optional brand blocks
    "dot_n1500": dict(seed=11, n=1500, q=12, d=64, sim="dot", batch_size=512),
    "dot_d768_n300": dict(seed=12, n=300, q=8, d=768, sim="dot", batch_size=128),
    "cos_block_n1300": dict(seed=13, n=1300, q=16, d=64, sim="cos", batch_size=512, brands=40,
                            queries_are_corpus=True, clustered=True),
    # heavy blocking: fewer than 1001 unmasked items for some rows -> -1e6 entries are returned
    "dot_block_tail_n1100": dict(seed=14, n=1100, q=10, d=64, sim="dot", batch_size=256, brands=5,
                                 queries_are_corpus=True),
}


# cases run through the UNMODIFIED reference ranking() with real .cuda() calls under
# torch.cuda.amp.autocast() on a B200 (tests/golden/make_golden_cuda.py)
CUDA_RANKING_CASES = {
    "dot_d768_n5000": dict(seed=21, n=5000, q=24, d=768, sim="dot", batch_size=512),
    "cos_block_d768_n4000": dict(seed=22, n=4000, q=20, d=768, sim="cos", batch_size=512, brands=30,
                                 queries_are_corpus=True, clustered=True),
    "dot_block_d64_n3000": dict(seed=23, n=3000, q=16, d=64, sim="dot", batch_size=256, brands=12,
                                queries_are_corpus=True),
}


def rime_case(name):
    """-> dict(U [B,d] f32, V [N,d] f32, prior csr float64 [B,N] or None, k)."""
    spec = RIME_CASES[name]
    b, n, d = spec["b"], spec["n"], spec["d"]
    U = embeddings(spec["seed"], b, d)
    V = embeddings(spec["seed"] + 100, n, d)
    prior = None
    if spec.get("prior"):
        rs = np.random.RandomState(spec["seed"] + 200)
        rows, cols, vals = [], [], []
        for r in range(b):
            # seen history: -1e10 (dataset/base.py:234); reranking candidates: +prior (:279-282)
            seen = rs.choice(n, size=rs.randint(0, 4), replace=False)
            cand = rs.choice(n, size=rs.randint(1, 6), replace=False)
            for c in seen:
                rows.append(r), cols.append(int(c)), vals.append(-1e10)
            for c in cand:
                rows.append(r), cols.append(int(c)), vals.append(float(spec["prior"]))
        prior = sps.csr_matrix((np.array(vals), (np.array(rows), np.array(cols))), shape=(b, n))
    return dict(U=U, V=V, prior=prior, k=spec["k"])


RIME_CASES = {
    "plain_k5": dict(seed=21, b=9, n=400, d=32, k=5),
    "mask_prior_k1": dict(seed=22, b=40, n=700, d=48, k=1, prior=1e5),
    "mask_prior_k10": dict(seed=23, b=17, n=257, d=64, k=10, prior=1.0),
}


def _zipf_text(rs, vocab, n_words):
    """Space-separated pseudo words w<i>, i ~ Zipf -> a few very common terms, a long tail."""
    ids = np.minimum(rs.zipf(1.3, size=n_words) - 1, vocab - 1)
    return " ".join(f"w{int(i)}" for i in ids)


def bm25_case(name):
    """-> dict(corpus {pid: text}, queries {qid: text}).  Texts are Zipfian pseudo-word
    sequences (sklearn's default token pattern keeps every ``w<i>`` token); a share of the
    queries are copies of corpus passages (long queries, >= 8 distinct terms, exercise numpy's
    pairwise row sum), some contain out-of-vocabulary words and repeated words, one is empty."""
    spec = BM25_CASES[name]
    rs = np.random.RandomState(spec["seed"])
    n, q, vocab = spec["n"], spec["q"], spec["vocab"]
    corpus = {f"p{i}": _zipf_text(rs, vocab, rs.randint(3, spec["doc_len"])) for i in range(n)}
    pids = list(corpus)
    queries = {}
    for i in range(q):
        kind = i % 4
        if kind == 0:
            text = corpus[pids[rs.randint(n)]]
        elif kind == 1:
            text = _zipf_text(rs, vocab, rs.randint(1, 6)) + " zzunseen" + str(i)
        elif kind == 2:
            w = _zipf_text(rs, vocab, 3)
            text = w + " " + w
        else:
            text = _zipf_text(rs, vocab, rs.randint(2, 12))
        queries[f"q{i}"] = text
    if spec.get("empty_query"):
        queries["q_empty"] = "zzunseen"
    return dict(corpus=corpus, queries=queries)


BM25_CASES = {
    # fewer than 1001 docs: the whole corpus is returned, zero-score tail in tie order
    "small_n600": dict(seed=31, n=600, q=12, vocab=300, doc_len=20, empty_query=True),
    # more than 1001 docs, many docs share no term with the query (zero scores below the cut)
    "tail_n3000": dict(seed=32, n=3000, q=16, vocab=2000, doc_len=40),
}


def al0_case(name):
    """Inputs of scripts/al_0_rank.py:136-218 (the request-building part): corpus / queries with
    characters the display filter drops, a dense and a BM25 ranking profile per query, the qid
    splits, STEP / N_REPEATS / REPEAT_SEED and (optionally) the landing-image table."""
    import pandas as pd

    spec = AL0_CASES[name]
    rs = np.random.RandomState(spec["seed"])
    n, q = spec["n"], spec["q"]
    junk = ["é", "—", "#", "@", "\"", "'", "%", "\n", "/", "*"]
    keep = list("abcXYZ019 ,:.;?$!()&[]")

    def text(length):
        chars = [junk[rs.randint(len(junk))] if rs.rand() < 0.15 else keep[rs.randint(len(keep))] for _ in range(length)]
        return "".join(chars)

    corpus = {f"d{i}": text(rs.randint(5, spec["text_len"])) for i in range(n)}
    same = spec.get("queries_are_corpus", False)
    queries = {pid: corpus[pid] for pid in list(corpus)[:q]} if same else {f"{1000 + i}": text(rs.randint(5, 60)) for i in range(q)}
    pids = np.array(list(corpus))
    depth = min(n, spec["depth"])

    def profile():
        out = {}
        for qid in queries:
            order = rs.permutation(n)[:depth]
            scores = np.sort(rs.standard_normal(depth).astype(np.float32))[::-1]
            out[qid] = dict(zip(pids[order].tolist(), scores.tolist()))
        return out

    dense, bm25 = profile(), profile()
    if spec.get("bm25_overlap"):  # BM25's best hits coincide with the dense top-2: the third comes later
        for qid in queries:
            head = list(dense[qid])[:2]
            rest = [p for p in bm25[qid] if p not in head]
            bm25[qid] = dict(zip(head + rest, sorted(bm25[qid].values(), reverse=True)))
    qids = list(queries)
    rs.shuffle(qids)
    nsplit = spec["splits"]
    qids_split = [qids[i::nsplit] for i in range(nsplit)]
    landing = None
    if spec.get("images"):
        landing = pd.Series({pid: f"https://img.example/{pid}.jpg" for pid in corpus})
    return dict(corpus=corpus, queries=queries, ranking_profile=dense, ranking_profile_bm25=bm25,
                qids_split=qids_split, number_of_qid_split_batch=nsplit, step=spec["step"],
                n_repeats=spec["n_repeats"], repeat_seed=spec["repeat_seed"], landing_image=landing)


AL0_CASES = {
    "nq_like_step1": dict(seed=51, n=400, q=60, depth=101, text_len=400, splits=4, step=1, n_repeats=3, repeat_seed=42,
                          bm25_overlap=True),
    # Prime-Pantry-like: queries are corpus items, images attached, tiny corpus so the random fill
    # collides with existing candidates (extra RNG draws), step wraps around the splits
    "pantry_like_step5": dict(seed=52, n=6, q=6, depth=3, text_len=300, splits=4, step=5, n_repeats=2,
                              repeat_seed=7, queries_are_corpus=True, images=True),
}


def al0_qrels(c):
    """qrels for an al0 case: for two queries in three the relevant passage is one of the dense
    top-3 or BM25 top-3 (so it is usually among the candidates), otherwise an unrelated one."""
    rs = np.random.RandomState(99)
    pids = list(c["corpus"])
    out = {}
    for n, qid in enumerate(c["queries"]):
        pool = list(c["ranking_profile"][qid])[:3] + list(c["ranking_profile_bm25"][qid])[:3]
        rel = pool[rs.randint(len(pool))] if n % 3 else pids[rs.randint(len(pids))]
        out[qid] = {rel: 1}
        if n % 5 == 0:
            out[qid][pool[0]] = 1  # two labelled candidates: the later one in shuffled order wins
    return out
