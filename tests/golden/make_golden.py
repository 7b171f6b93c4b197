#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the UNMODIFIED reference code in this container.

    python tests/golden/make_golden.py

Needs /root/reference (absent on the GPU box -> the committed .npz files are what travels).
Each golden stores only the reference's outputs; inputs are regenerated from seeds by
tests/golden/cases.py.

* ranking_<case>.npz : output of scripts/ms_marco_eval.py::ranking (real function, with
  Tensor.cuda()/torch.cuda.synchronize patched to no-ops because this box has no GPU):
  per query the ordered corpus positions and float scores (<=1001 entries).
* bm25_<case>.npz    : outputs of the real scripts/bm_25.py::BM25 (fit + transform: float64 score
  vectors of the first queries) and of scripts/ms_marco_eval.py::ranking_bm25 (ordered positions
  and float scores, <=1001 entries per query).
* al0_<case>.npz     : bytes of request_orig.csv / request_perm.csv and the id_track dict written by
  the request-building statements of scripts/al_0_rank.py ("## creation" to the end of the file,
  :136-218).  The script is module-level code that cannot be imported (argparse, downloads, an
  ImportError at :28-34), so exactly those statements are read from the reference file at
  generation time and executed unmodified in a namespace holding the synthetic inputs.
* rime_<case>.npz    : outputs of the real rime_lite code: ``_assign_topk`` CSR indices,
  the dense ``as_tensor`` matrix dtype, ``_argsort`` head, ``evaluate_item_rec`` metrics.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import cases  # noqa: E402
import _ref_loader  # noqa: E402


def make_ranking():
    me = _ref_loader.load_ms_marco_eval()
    for name in cases.RANKING_CASES:
        c = cases.ranking_case(name)
        os.environ["CCREC_SIM_TYPE"] = c["sim_type"]
        with _ref_loader.cpu_as_cuda():
            prof = me.ranking(c["corpus"], c["queries"], cases.TextTable(c["table"]), c["batch_size"],
                              c["block_dict"])
        qids = list(c["queries"].keys())
        pos = {pid: i for i, pid in enumerate(c["corpus"].keys())}
        order = np.array([[pos[p] for p in prof[q].keys()] for q in qids], dtype=np.int32)
        scores = np.array([list(prof[q].values()) for q in qids], dtype=np.float64)
        np.savez_compressed(os.path.join(HERE, f"ranking_{name}.npz"), order=order, scores=scores)
        print("ranking", name, order.shape)


def make_rime():
    ru = _ref_loader.load_rime_lite_util()
    rm = _ref_loader.load_rime_lite_metrics()
    for name in cases.RIME_CASES:
        c = cases.rime_case(name)
        S = ru.LazyDenseMatrix(c["U"]) @ ru.LazyDenseMatrix(c["V"]).T
        if c["prior"] is not None:
            S = S + c["prior"]
        dense = S.as_tensor("cpu")
        torch.manual_seed(0)
        csr = ru._assign_topk(S, c["k"], device="cpu")
        idx_topk_order = csr.indices.reshape(len(c["U"]), c["k"]).astype(np.int32).copy()
        # batch_size=4 exercises the row-slicing / collate contract (util/__init__.py:126-131)
        torch.manual_seed(0)
        csr_b = ru._assign_topk(S, c["k"], device="cpu", batch_size=4)
        idx_b_topk_order = csr_b.indices.reshape(len(c["U"]), c["k"]).astype(np.int32).copy()
        ar, ac = ru._argsort(S, tie_breaker=0, device="cpu")
        target = (csr_b > 0).astype(np.float64)  # any 0/1 target works for the metric replay
        torch.manual_seed(0)
        metrics = rm.evaluate_item_rec(target, S, c["k"])
        np.savez_compressed(
            os.path.join(HERE, f"rime_{name}.npz"),
            indices=idx_topk_order,  # copied before scipy comparisons sort them in place
            indices_batched=idx_b_topk_order,
            indptr=csr.indptr.astype(np.int64),
            data=csr.data,
            shape=np.array(csr.shape),
            dense_dtype=str(dense.dtype),
            dense_head=dense[:4, :16].numpy().astype(np.float64),
            argsort_rows=ar[:64].astype(np.int32),
            argsort_cols=ac[:64].astype(np.int32),
            metric_names=np.array(sorted(metrics)),
            metric_values=np.array([metrics[m] for m in sorted(metrics)], dtype=np.float64),
        )
        print("rime", name, csr.shape, dense.dtype, metrics)


def make_bm25():
    me = _ref_loader.load_ms_marco_eval()
    import bm_25  # scripts/bm_25.py, on sys.path after load_ms_marco_eval

    for name in cases.BM25_CASES:
        c = cases.bm25_case(name)
        prof = me.ranking_bm25(c["corpus"], c["queries"])
        qids = list(c["queries"].keys())
        pos = {pid: i for i, pid in enumerate(c["corpus"].keys())}
        order = np.array([[pos[p] for p in prof[q].keys()] for q in qids], dtype=np.int32)
        scores = np.array([list(prof[q].values()) for q in qids], dtype=np.float64)
        model = bm_25.BM25(b=0.75, k1=1.2).fit(list(c["corpus"].values()))
        dense = np.stack([model.transform(c["queries"][q]) for q in qids[:8]])
        model16 = bm_25.BM25().fit(list(c["corpus"].values()))  # class defaults b=0.75, k1=1.6
        dense16 = np.stack([model16.transform(c["queries"][q]) for q in qids[:4]])
        np.savez_compressed(os.path.join(HERE, f"bm25_{name}.npz"), order=order, scores=scores, dense=dense,
                            dense_k16=dense16, avdl=np.float64(model.avdl), vocab=len(model.vectorizer.vocabulary_))
        print("bm25", name, order.shape, dense.shape, "nonzero/row", (dense > 0).sum(1))


def make_al0():
    import re
    import tempfile

    import pandas as pd

    src = open(os.path.join(_ref_loader.REFERENCE_ROOT, "scripts", "al_0_rank.py")).read()
    creation = src[src.index("## creation"):]
    os.environ["CCREC_DISPLAY_LENGTH"] = "250"  # al_0_rank.py:13
    for name in cases.AL0_CASES:
        c = cases.al0_case(name)
        with tempfile.TemporaryDirectory() as tmp:
            ns = dict(np=np, pd=pd, re=re, os=os, torch=torch, STEP=c["step"], corpus=c["corpus"], queries=c["queries"],
                      ranking_profile=c["ranking_profile"], ranking_profile_bm25=c["ranking_profile_bm25"],
                      qids_split=c["qids_split"], number_of_qid_split_batch=c["number_of_qid_split_batch"],
                      landingImage=c["landing_image"], current_working_dir=tmp, N_REPEATS=c["n_repeats"],
                      REPEAT_SEED=c["repeat_seed"])
            exec(compile(creation, "al_0_rank.py[creation]", "exec"), ns)
            orig = open(os.path.join(tmp, "request_orig.csv"), "rb").read()
            perm = open(os.path.join(tmp, "request_perm.csv"), "rb").read()
            track = torch.load(os.path.join(tmp, "id_track.pt"))
        # scripts/al_oracle_agent.py:134-180 (generate_train_data): the module cannot be imported
        # (encoder / lightning imports, module-level work), so the function's own source text is executed
        import random

        agent = open(os.path.join(_ref_loader.REFERENCE_ROOT, "scripts", "al_oracle_agent.py")).read()
        fn_src = agent[agent.index("def generate_train_data("):agent.index("def combine_train_data(")]
        ns2 = dict(np=np, random=random)
        exec(compile(fn_src, "al_oracle_agent.py[generate_train_data]", "exec"), ns2)
        qrels = cases.al0_qrels(c)
        qids = c["qids_split"][c["step"] % c["number_of_qid_split_batch"]]
        train = {}
        for variant, keys in (("plain", []), ("attention", list(c["corpus"].keys()))):
            random.seed(1234)
            train[variant] = ns2["generate_train_data"](qids, qrels, c["ranking_profile"], c["ranking_profile_bm25"],
                                                        keys, c["step"])
        import json

        np.savez_compressed(os.path.join(HERE, f"al0_{name}.npz"), request_orig=np.frombuffer(orig, dtype=np.uint8),
                            train_data=np.array(json.dumps(train, sort_keys=True)),
                            request_perm=np.frombuffer(perm, dtype=np.uint8),
                            id_track_keys=np.array(list(track.keys()), dtype=object).astype(str),
                            id_track_vals=np.array(list(track.values()), dtype=object).astype(str))
        print("al0", name, len(orig), len(perm), len(track))


if __name__ == "__main__":
    assert _ref_loader.reference_available(), "needs /root/reference"
    what = sys.argv[1:] or ["ranking", "rime", "bm25", "al0"]
    if "ranking" in what:
        make_ranking()
    if "rime" in what:
        make_rime()
    if "bm25" in what:
        make_bm25()
    if "al0" in what:
        make_al0()
