"""al_0_rank request building (scripts/al_0_rank.py:136-218): the oracle restatement and the
product's host logic (ccr_b200.al_rank) against outputs of the reference's own statements
(tests/golden/al0_*.npz).  Pure host code: runs without a GPU."""
import io
import os

import numpy as np
import pandas as pd
import pytest
import torch

import cases
from oracle import ccr_oracle as O


def _csv(header, rows):
    buf = io.StringIO()
    pd.DataFrame(rows, columns=header).to_csv(buf, index=False)
    return buf.getvalue().encode()


@pytest.mark.parametrize("name", list(cases.AL0_CASES))
def test_oracle_requests_match_reference(name, golden_dir):
    g = np.load(os.path.join(golden_dir, f"al0_{name}.npz"))
    c = cases.al0_case(name)
    header, rows, perm, track = O.al0_requests_ref(
        c["ranking_profile"], c["ranking_profile_bm25"], c["corpus"], c["queries"], c["qids_split"], c["step"],
        c["number_of_qid_split_batch"], c["n_repeats"], c["repeat_seed"], c["landing_image"])
    assert _csv(header, rows) == g["request_orig"].tobytes()
    assert _csv(header, perm) == g["request_perm"].tobytes()
    assert list(track.keys()) == g["id_track_keys"].tolist() and list(track.values()) == g["id_track_vals"].tolist()


@pytest.mark.parametrize("name", list(cases.AL0_CASES))
def test_product_requests_match_reference(name, golden_dir, tmp_path, monkeypatch):
    """Same ranking_profile in -> byte-identical request_orig.csv / request_perm.csv / id_track out."""
    from ccr_b200 import al_rank

    monkeypatch.setenv("CCREC_DISPLAY_LENGTH", "250")
    g = np.load(os.path.join(golden_dir, f"al0_{name}.npz"))
    c = cases.al0_case(name)
    split = c["qids_split"][c["step"] % c["number_of_qid_split_batch"]]
    header, rows, track = al_rank.build_requests(c["ranking_profile"], c["ranking_profile_bm25"], c["corpus"],
                                                 c["queries"], split, c["step"], c["landing_image"])
    al_rank.write_requests(str(tmp_path), header, rows, track, c["n_repeats"], c["repeat_seed"])
    assert open(tmp_path / "request_orig.csv", "rb").read() == g["request_orig"].tobytes()
    assert open(tmp_path / "request_perm.csv", "rb").read() == g["request_perm"].tobytes()
    saved = torch.load(tmp_path / "id_track.pt")
    assert list(saved.keys()) == g["id_track_keys"].tolist() and list(saved.values()) == g["id_track_vals"].tolist()


def test_rank_step_reuses_cached_profile(tmp_path, monkeypatch):
    """al_0_rank.py:117-118: an existing ranking_profile.pt is loaded instead of recomputed -- so the
    whole step runs without touching the device; outputs land in data_iteration_{STEP}/."""
    from ccr_b200 import al_rank

    monkeypatch.setenv("CCREC_DISPLAY_LENGTH", "250")
    c = cases.al0_case("nq_like_step1")
    wd = tmp_path / f"data_iteration_{c['step']}"
    wd.mkdir()
    torch.save(c["ranking_profile"], wd / "ranking_profile.pt")
    qid0 = next(iter(c["queries"]))
    qrels = {qid0: {next(iter(c["ranking_profile"][qid0])): 1}}

    def boom(_texts):
        raise AssertionError("embedding_func must not be called when the profile is cached")

    prof, mrr, orig, perm = al_rank.rank_step(c["corpus"], c["queries"], qrels, boom, str(tmp_path), c["step"],
                                              c["ranking_profile_bm25"], c["qids_split"], c["n_repeats"],
                                              c["repeat_seed"])
    assert prof == c["ranking_profile"] and mrr["MRR@1"] == 1.0
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "al0_nq_like_step1.npz"))
    assert open(wd / "request_perm.csv", "rb").read() == g["request_perm"].tobytes()
    assert len(perm) == c["n_repeats"] * len(orig)


def test_mrr_from_profile_follows_beir_semantics():
    from ccr_b200 import al_rank

    prof = {"q0": {"a": 3.0, "b": 2.0, "c": 1.0}, "q1": {"a": 0.1, "c": 0.9, "b": 0.5}, "q2": {"a": 1.0}}
    qrels = {"q0": {"b": 1}, "q1": {"a": 2, "b": 0}, "q2": {"z": 1}, "q3": {"a": 1}}
    m = al_rank.mrr_from_profile(qrels, prof, (1, 2, 3))
    # q0: first hit at rank 2; q1: ordered c, b, a -> rank 3; q2: none; divided by len(qrels) = 4
    assert m == {"MRR@1": 0.0, "MRR@2": round(0.5 / 4, 5), "MRR@3": round((0.5 + 1 / 3) / 4, 5)}


@pytest.mark.parametrize("name", list(cases.AL0_CASES))
def test_generate_train_data_matches_reference(name, golden_dir):
    """scripts/al_oracle_agent.py:134-180 (function source executed for the golden) vs the drop-in:
    same candidates, same shuffle (global ``random``, seeded here), same positive / negative split."""
    import json
    import random

    from ccr_b200 import al_rank

    g = np.load(os.path.join(golden_dir, f"al0_{name}.npz"))
    want = json.loads(str(g["train_data"]))
    c = cases.al0_case(name)
    qrels = cases.al0_qrels(c)
    qids = c["qids_split"][c["step"] % c["number_of_qid_split_batch"]]
    for variant, keys in (("plain", []), ("attention", list(c["corpus"].keys()))):
        random.seed(1234)
        got = al_rank.generate_train_data(qids, qrels, c["ranking_profile"], c["ranking_profile_bm25"], keys, c["step"])
        assert got == want[variant], variant
        assert list(got.keys()) == [q for q in qids if q in want[variant]]


def test_block_drawn_candidates_follow_the_scalar_choice_stream():
    """The vectorised random fill must consume the ``RandomState(STEP).choice(n)`` stream exactly like the
    reference's one-call-per-attempt loop (al_0_rank.py:178-182): tiny corpora force many rejections."""
    from ccr_b200 import al_rank

    for n, Q, seed in [(5, 300, 0), (8, 1000, 1), (1500, 700, 2), (9862, 50, 3)]:
        a, b = np.random.RandomState(seed), np.random.RandomState(seed)
        np.testing.assert_array_equal([a.choice(n) for _ in range(500)], b.randint(0, n, size=500))
        rs = np.random.RandomState(seed + 10)
        cand_pos = np.array([rs.choice(n, size=3, replace=False) for _ in range(Q)], dtype=np.int64)
        want = []
        ref = np.random.RandomState(seed)
        for c in cand_pos:
            while True:
                d = ref.choice(n)
                if d not in c:
                    want.append(d)
                    break
        got = al_rank._fill_fourth(cand_pos, al_rank._Draws(np.random.RandomState(seed), n, block=64))
        np.testing.assert_array_equal(got, want)


def test_ranking_profile_behaves_like_the_reference_dict(tmp_path):
    from ccr_b200.ranking import RankingProfile

    order = np.array([[2, 0, 1], [1, 2, 0]])
    scores = np.array([[3.0, 2.0, -1e6], [9.0, 8.5, 1.0]], dtype=np.float32)
    prof = RankingProfile(["q0", "q1"], ["a", "b", "c"], scores, order)
    plain = {"q0": {"c": 3.0, "a": 2.0, "b": -1e6}, "q1": {"b": 9.0, "c": 8.5, "a": 1.0}}
    assert list(prof) == ["q0", "q1"] and len(prof) == 2 and "q1" in prof and "zz" not in prof
    assert list(prof["q0"].items()) == list(plain["q0"].items())      # descending order preserved
    assert prof == plain and plain == prof and dict(prof.items()) == plain
    assert prof.top_ids("q1", 2) == ["b", "c"]
    with pytest.raises(KeyError):
        prof["zz"]
    import pickle

    path = tmp_path / "ranking_profile.pt"
    torch.save(prof.to_dict(), path)                                   # what rank_step writes
    back = torch.load(path)
    assert type(back) is dict and back == plain                        # the reference's file format
    again = pickle.loads(pickle.dumps(prof))                           # generic pickling also yields the dict
    assert type(again) is dict and again == plain
