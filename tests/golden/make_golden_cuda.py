#!/usr/bin/env python3
"""Golden vectors of the reference's REAL GPU path: the unmodified
scripts/ms_marco_eval.py::ranking with its own .cuda() / torch.cuda.synchronize() calls, wrapped in
torch.cuda.amp.autocast() exactly as scripts/al_0_rank.py:125 calls it, run on a B200.

/root/reference does not exist on the GPU box, so the reference tree is staged (unmodified, by
stage_reference.sh) under baseline/_ref/reference -- git-ignored, never part of the repository -- and
this script is run there once:

    bash tests/golden/stage_reference.sh                       # build container
    gpurun -- 'CCR_REFERENCE_ROOT=baseline/_ref/reference python tests/golden/make_golden_cuda.py gpurun_out/'
    cp gpurun_out/ranking_cuda_autocast_*.npz tests/golden/    # commit the outputs

Each .npz stores only the reference's outputs (ordered corpus positions, float scores, <= 1001 per
query); inputs are regenerated from seeds by cases.py.  tests replay them against the oracle's
restatement of the fp16-autocast arithmetic (CPU) and against the CUDA product (GPU).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import cases  # noqa: E402
import _ref_loader  # noqa: E402

out_dir = sys.argv[1] if len(sys.argv) > 1 else HERE
assert torch.cuda.is_available(), "this generator needs the GPU: it records the reference's CUDA arithmetic"
assert _ref_loader.reference_available(), f"no reference tree at {_ref_loader.REFERENCE_ROOT}"
me = _ref_loader.load_ms_marco_eval()
os.makedirs(out_dir, exist_ok=True)
for name in cases.CUDA_RANKING_CASES:
    c = cases.ranking_case(name)
    os.environ["CCREC_SIM_TYPE"] = c["sim_type"]
    with torch.cuda.amp.autocast():
        prof = me.ranking(c["corpus"], c["queries"], cases.TextTable(c["table"]), c["batch_size"], c["block_dict"])
    qids = list(c["queries"].keys())
    pos = {pid: i for i, pid in enumerate(c["corpus"].keys())}
    order = np.array([[pos[p] for p in prof[q].keys()] for q in qids], dtype=np.int32)
    scores = np.array([list(prof[q].values()) for q in qids], dtype=np.float64)
    np.savez_compressed(os.path.join(out_dir, f"ranking_cuda_autocast_{name}.npz"), order=order, scores=scores,
                        device=torch.cuda.get_device_name(0), torch=torch.__version__)
    print("ranking (cuda, autocast)", name, order.shape, "distinct fp16 scores in row 0:", len(set(scores[0])))
