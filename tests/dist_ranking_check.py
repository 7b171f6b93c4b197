#!/usr/bin/env python3
"""ranking_sharded on real GPUs (torchrun, NCCL) == the single-table ranking() on the same inputs:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tests/dist_ranking_check.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "crowd-coachable-recommendations_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import cases  # noqa: E402
import ccr_b200  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
ok = True
for name in ("dot_block_tail_n1100", "cos_block_n1300", "dot_n1500"):
    c = cases.ranking_case(name)
    os.environ["CCREC_SIM_TYPE"] = c["sim_type"]
    one = ccr_b200.ranking(c["corpus"], c["queries"], cases.TextTable(c["table"]), c["batch_size"], c["block_dict"], device=dev)
    table = cases.TextTable(c["table"])
    many = ccr_b200.ranking_sharded(c["corpus"], c["queries"], table, c["batch_size"], c["block_dict"], device=dev)
    same = list(one) == list(many) and all(list(one[q].items()) == list(many[q].items()) for q in one)
    print(f"rank {dist.get_rank()} {name}: identical={same} encoder_calls={table.calls}", flush=True)
    ok &= same
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
