"""Pin the oracle (oracle/ccr_oracle.py) against outputs of the UNMODIFIED reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py from /root/reference)."""
import os

import numpy as np
import pytest
import torch

import cases
from oracle import ccr_oracle as O


def _profile_to_arrays(prof, c):
    pos = {pid: i for i, pid in enumerate(c["corpus"].keys())}
    qids = list(c["queries"].keys())
    order = np.array([[pos[p] for p in prof[q].keys()] for q in qids])
    scores = np.array([list(prof[q].values()) for q in qids])
    return order, scores


@pytest.mark.parametrize("name", list(cases.RANKING_CASES))
def test_ranking_ref_matches_reference(name, golden_dir):
    g = np.load(os.path.join(golden_dir, f"ranking_{name}.npz"))
    c = cases.ranking_case(name)
    prof = O.ranking_ref(c["corpus"], c["queries"], cases.TextTable(c["table"]), c["batch_size"],
                         c["block_dict"], sim_type=c["sim_type"])
    order, scores = _profile_to_arrays(prof, c)
    assert order.shape == g["order"].shape
    np.testing.assert_array_equal(scores, g["scores"])  # same torch ops -> bit-identical
    # order may differ only inside exact ties (the reference's sort is unstable): the -1e6
    # tail of heavily blocked rows may hold ANY subset of the blocked ids, in any order.
    pos = {pid: i for i, pid in enumerate(c["corpus"].keys())}
    for b, qid in enumerate(c["queries"].keys()):
        live = scores[b] > -1e6
        if not np.array_equal(order[b][live], g["order"][b][live]):
            # exact fp32 ties may be permuted: same ids within each run of equal scores
            sl = scores[b][live]
            for v in np.unique(sl[order[b][live] != g["order"][b][live]]):
                run = sl == v
                assert set(order[b][live][run]) == set(g["order"][b][live][run])
        if (~live).any():
            blocked = {pos[p] for p in c["block_dict"][qid]}
            assert set(order[b][~live]) <= blocked and set(g["order"][b][~live]) <= blocked
            assert len(set(order[b])) == order.shape[1]


@pytest.mark.parametrize("name", list(cases.CUDA_RANKING_CASES))
def test_ranking_ref_matches_reference_cuda_autocast_path(name, golden_dir):
    """the goldens are outputs of the UNMODIFIED ms_marco_eval.ranking run on a
    B200 with its real .cuda() calls inside torch.cuda.amp.autocast() (al_0_rank.py:125), i.e. fp16
    tensor-core scores.  The oracle's CPU restatement of that arithmetic must reproduce them up to one
    fp16 ulp of cuBLAS's accumulation order (and the permutations that such a one-ulp move, or an exact
    fp16 tie under the reference's unstable sort, allows)."""
    g = np.load(os.path.join(golden_dir, f"ranking_cuda_autocast_{name}.npz"))
    c = cases.ranking_case(name)
    prof = O.ranking_ref(c["corpus"], c["queries"], cases.TextTable(c["table"]), c["batch_size"], c["block_dict"],
                         sim_type=c["sim_type"], autocast_fp16=True)
    order, scores = _profile_to_arrays(prof, c)
    assert order.shape == g["order"].shape
    live = g["scores"] > -1e6
    assert np.array_equal(live, scores > -1e6)
    ulp = np.abs(g["scores"]) * 2.0 ** -10 + 2.0 ** -24      # one fp16 ulp at the score's binade (or coarser)
    assert (np.abs(scores - g["scores"])[live] <= ulp[live]).all()
    assert (scores == g["scores"]).mean() > 0.99
    assert (order == g["order"])[live].mean() > 0.99
    errs = O.check_topk(scores, order, ref_scores=g["scores"], ref_ids=g["order"], rtol=2.0 ** -9, atol=1e-6)
    assert not errs, errs[:3]


@pytest.mark.parametrize("name", list(cases.RIME_CASES))
def test_rime_ref_matches_reference(name, golden_dir):
    g = np.load(os.path.join(golden_dir, f"rime_{name}.npz"))
    c = cases.rime_case(name)
    dense = O.lazy_score_dense_ref(c["U"], c["V"], c["prior"])
    assert str(dense.dtype) == str(g["dense_dtype"])
    np.testing.assert_allclose(dense[:4, :16].numpy(), g["dense_head"], rtol=0, atol=0)
    csr = O.assign_topk_ref(dense, c["k"])
    got = csr.indices.reshape(len(c["U"]), c["k"])
    np.testing.assert_array_equal(got, g["indices"])           # no exact ties in these cases
    np.testing.assert_array_equal(got, g["indices_batched"])
    np.testing.assert_array_equal(csr.indptr, g["indptr"])
    np.testing.assert_array_equal(csr.data, g["data"])
    ar, ac = O.argsort_ref(dense.numpy())
    np.testing.assert_array_equal(ar[:64], g["argsort_rows"])
    np.testing.assert_array_equal(ac[:64], g["argsort_cols"])


@pytest.mark.parametrize("name", list(cases.RIME_CASES))
def test_score_topk_ref_agrees_with_rime_semantics(name):
    """The kernel arbiter (fp32 inputs, no bf16 rounding) reproduces the reference's top-k."""
    c = cases.rime_case(name)
    dense = O.lazy_score_dense_ref(c["U"], c["V"], c["prior"])
    want = O.assign_topk_ref(dense, c["k"]).indices.reshape(len(c["U"]), c["k"])
    mask, mode = None, O.MASK_NONE
    if c["prior"] is not None:
        p = c["prior"].tocsr()
        mask, mode = (p.indptr, p.indices, p.data), O.MASK_ADD
    s, i = O.score_topk_ref(c["U"], c["V"], c["k"], mask=mask, mode=mode, round_bf16=False, chunk=97)
    np.testing.assert_array_equal(i.numpy(), want)


def test_score_topk_ref_set_mode_matches_ranking():
    c = cases.ranking_case("dot_block_tail_n1100")
    prof = O.ranking_ref(c["corpus"], c["queries"], cases.TextTable(c["table"]), c["batch_size"],
                         c["block_dict"], sim_type="dot")
    order, scores = _profile_to_arrays(prof, c)
    n = len(c["corpus"])
    q = len(c["queries"])
    pos = {pid: i for i, pid in enumerate(c["corpus"].keys())}
    indptr, cols = [0], []
    for qid in c["queries"]:
        cols += [pos[p] for p in c["block_dict"][qid]]
        indptr.append(len(cols))
    vals = np.full(len(cols), -1e6)
    Q = c["table"][:q]
    s, i = O.score_topk_ref(Q, c["table"][:n], 1001, mask=(indptr, cols, vals), mode=O.MASK_SET,
                            round_bf16=False, chunk=300)
    np.testing.assert_allclose(s.numpy(), scores, rtol=1e-6)
    np.testing.assert_array_equal(i.numpy(), order)  # both stable: ties -> lowest position


def test_check_topk_rule():
    rs = np.random.RandomState(0)
    full = rs.standard_normal((3, 50))
    order = np.argsort(-full, axis=1)[:, :5]
    sc = np.take_along_axis(full, order, 1)
    assert O.check_topk(sc, order, full_scores=full) == []
    bad = order.copy()
    bad[0, 4] = np.argsort(-full[0])[30]
    sc_bad = np.take_along_axis(full, bad, 1)
    assert O.check_topk(sc_bad, bad, full_scores=full)
    assert O.check_topk(sc, order, ref_scores=sc, ref_ids=order) == []
    with pytest.raises(RuntimeError):
        O.score_topk_ref(torch.zeros(2, 8), torch.zeros(3, 8), 4)


@pytest.mark.parametrize("name", list(cases.BM25_CASES))
def test_bm25_ref_matches_reference(name, golden_dir):
    """BM25Ref / ranking_bm25_ref vs the real scripts/bm_25.py + ranking_bm25 outputs."""
    g = np.load(os.path.join(golden_dir, f"bm25_{name}.npz"))
    c = cases.bm25_case(name)
    texts = list(c["corpus"].values())
    qids = list(c["queries"].keys())
    model = O.BM25Ref(b=0.75, k1=1.2).fit(texts)
    assert float(model.avdl) == float(g["avdl"]) and len(model.vectorizer.vocabulary_) == int(g["vocab"])
    dense = np.stack([model.transform(c["queries"][q]) for q in qids[: len(g["dense"])]])
    np.testing.assert_array_equal(dense, g["dense"])  # same float64 arithmetic -> bit-identical
    model16 = O.BM25Ref().fit(texts)
    dense16 = np.stack([model16.transform(c["queries"][q]) for q in qids[: len(g["dense_k16"])]])
    np.testing.assert_array_equal(dense16, g["dense_k16"])
    prof = O.ranking_bm25_ref(c["corpus"], c["queries"])
    pos = {pid: i for i, pid in enumerate(c["corpus"].keys())}
    order = np.array([[pos[p] for p in prof[q].keys()] for q in qids])
    scores = np.array([list(prof[q].values()) for q in qids])
    assert order.shape == g["order"].shape
    np.testing.assert_array_equal(scores, g["scores"])
    for b in range(len(qids)):  # ids may differ only inside runs of exactly tied float32 scores
        if not np.array_equal(order[b], g["order"][b]):
            diff = order[b] != g["order"][b]
            for v in np.unique(scores[b][diff]):
                run = scores[b] == v
                last = np.nonzero(run)[0][-1] == order.shape[1] - 1  # a run cut by the 1001 slice
                if not last:
                    assert set(order[b][run]) == set(g["order"][b][run])
