#!/usr/bin/env python3
"""Diagnostic sweep for the GPU box (not collected by pytest).

    python tests/gpu_probe.py all            # every case, each in its own subprocess + timeout
    python tests/gpu_probe.py case <json>    # one case in this process

One JSON line per case goes to stdout and gpurun_out/probe.jsonl: parity vs a torch fp32
reference on the same bf16-rounded inputs, plus CUDA-event timing.  A kernel trap / hang in one
case cannot take the others down.
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "crowd-coachable-recommendations_b200"))

CASES = [
    # name, algo(1 simt, 2 tc), B, N, D, k, mask(0/1/2), timing iters
    dict(name="simt_tiny", algo=1, B=3, N=1000, D=64, k=10, mask=0),
    dict(name="simt_full_k", algo=1, B=8, N=600, D=768, k=600, mask=0),
    dict(name="simt_mask_set", algo=1, B=5, N=5000, D=768, k=100, mask=1),
    dict(name="simt_mask_add", algo=1, B=20, N=3000, D=128, k=7, mask=2),
    dict(name="simt_k1001", algo=1, B=4, N=20000, D=768, k=1001, mask=0),
    dict(name="tc_gemm_full", algo=2, B=128, N=512, D=768, k=512, mask=0),
    dict(name="tc_gemm_ragged", algo=2, B=77, N=1000, D=200, k=1000, mask=0),
    dict(name="tc_small", algo=2, B=256, N=100000, D=768, k=100, mask=0),
    dict(name="tc_mask_set", algo=2, B=130, N=50000, D=768, k=100, mask=1),
    dict(name="tc_mask_add", algo=2, B=64, N=30000, D=768, k=10, mask=2),
    dict(name="tc_k1001", algo=2, B=200, N=200000, D=768, k=1001, mask=0),
    dict(name="tc_nq", algo=2, B=512, N=2681468, D=768, k=100, mask=0, iters=3),
    dict(name="simt_b1_8m", algo=1, B=1, N=8841823, D=768, k=100, mask=0, iters=3),
    dict(name="simt_b8_8m", algo=1, B=8, N=8841823, D=768, k=100, mask=0, iters=3),
    dict(name="tc_b16_8m", algo=2, B=16, N=8841823, D=768, k=100, mask=0, iters=3),
    dict(name="tc_b128_8m", algo=2, B=128, N=8841823, D=768, k=100, mask=0, iters=3),
    dict(name="tc_b1024_8m", algo=2, B=1024, N=8841823, D=768, k=100, mask=1, iters=3),
    dict(name="tc_b4096_8m", algo=2, B=4096, N=8841823, D=768, k=100, mask=1, iters=3, check_rows=64),
]


def run_case(c):
    import numpy as np
    import torch

    import ccr_b200
    from ccr_b200 import engine

    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1234)
    B, N, D, k = c["B"], c["N"], c["D"], c["k"]
    items = torch.empty((N, D), dtype=torch.bfloat16, device=dev)
    for s in range(0, N, 1 << 20):
        e = min(N, s + (1 << 20))
        items[s:e] = torch.randn((e - s, D), generator=g, device=dev).to(torch.bfloat16)
    q = torch.randn((B, D), generator=g, device=dev).to(torch.bfloat16)
    mask = None
    if c["mask"]:
        rs = np.random.RandomState(5)
        rows = [np.unique(rs.randint(0, N, size=min(N, rs.randint(0, 65)))) for _ in range(B)]
        if c["mask"] == 1:
            mask = engine.SparseMask.from_lists(rows, N, -1e6, engine.MASK_SET, dev)
        else:
            indptr = np.concatenate([[0], np.cumsum([len(r) for r in rows])])
            cols = np.concatenate([np.sort(r) for r in rows]) if rows else np.zeros(0)
            vals = np.where(rs.rand(len(cols)) < 0.5, -1e10, 1e5)
            mask = engine.SparseMask(indptr, cols, vals, N, engine.MASK_ADD, dev)
    out = {"name": c["name"], "B": B, "N": N, "D": D, "k": k, "algo": c["algo"], "mask": c["mask"]}
    out["plan"] = ccr_b200._lib.plan_info(B, N, D, k, c["algo"])
    t0 = time.time()
    sc, ids, sc64 = engine.score_topk(q, items, k, mask=mask, algo=c["algo"], want_f64=True)
    torch.cuda.synchronize()
    out["first_call_s"] = round(time.time() - t0, 4)

    # reference on sampled rows: fp32 scores of the bf16 inputs, full row
    check_rows = min(B, c.get("check_rows", 32))
    ridx = torch.linspace(0, B - 1, check_rows).long().unique()
    qf = q[ridx].float()
    ref = torch.empty((len(ridx), N), dtype=torch.float64 if c["mask"] == 2 else torch.float32, device=dev)
    for s in range(0, N, 1 << 20):
        e = min(N, s + (1 << 20))
        ref[:, s:e] = (qf @ items[s:e].float().T).to(ref.dtype)
    if mask is not None:
        ip, cc, vv = mask.host
        for j, r in enumerate(ridx.tolist()):
            cols = torch.as_tensor(cc[ip[r]:ip[r + 1]].astype(np.int64), device=dev)
            vals = torch.as_tensor(vv[ip[r]:ip[r + 1]], device=dev).to(ref.dtype)
            if c["mask"] == 1:
                ref[j, cols] = vals
            else:
                ref[j, cols] += vals
    if N > 2_000_000:
        rs_, ri_ = torch.topk(ref, k, dim=1)
    else:
        rs_, ri_ = torch.sort(ref, dim=1, descending=True, stable=True)
        rs_, ri_ = rs_[:, :k], ri_[:, :k]
    got_i = ids[ridx]
    got_s = sc64[ridx]
    # score of every returned id under the reference
    true = torch.gather(ref, 1, got_i.clamp(min=0)).double()
    rel = ((got_s - true).abs() / true.abs().clamp(min=1e-3)).max().item()
    kth = rs_[:, -1:].double()
    tol = 1e-2 * kth.abs() + 1e-6
    below = (true < kth - tol).sum().item()
    exact_ids = (got_i == ri_).float().mean().item()
    set_ok = 0
    for j in range(len(ridx)):
        set_ok += int(set(got_i[j].tolist()) == set(ri_[j].tolist()))
    desc = bool((got_s[:, 1:] <= got_s[:, :-1]).all().item())
    dup = int(sum(len(set(r.tolist())) != k for r in got_i))
    out.update(max_rel_err=rel, ids_below_kth=below, ids_exact_frac=exact_ids, rows_set_equal=set_ok,
               rows_checked=len(ridx), descending=desc, rows_with_dups=dup,
               ok=bool(rel < 1e-2 and below == 0 and desc and dup == 0 and set_ok >= len(ridx) - 1))
    iters = c.get("iters", 0)
    if iters:
        for _ in range(2):
            engine.score_topk(q, items, k, mask=mask, algo=c["algo"])
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(iters):
            engine.score_topk(q, items, k, mask=mask, algo=c["algo"])
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / iters
        out["ms"] = round(ms, 3)
        out["qps"] = round(B / ms * 1e3, 1)
        out["table_GBps"] = round(N * D * 2 / ms / 1e6, 1)
        out["TFLOPs"] = round(2.0 * B * N * D / ms / 1e9, 1)
    return out


def main():
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = os.path.join(ROOT, "gpurun_out", "probe.jsonl")
    if sys.argv[1] == "case":
        c = json.loads(sys.argv[2])
        try:
            res = run_case(c)
        except Exception as e:  # noqa: BLE001
            res = {"name": c["name"], "ok": False, "error": f"{type(e).__name__}: {e}"[:500]}
        print("PROBE " + json.dumps(res), flush=True)
        return
    only = sys.argv[2].split(",") if len(sys.argv) > 2 else None
    for c in CASES:
        if only and not any(c["name"].startswith(o) for o in only):
            continue
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "case", json.dumps(c)],
                               capture_output=True, text=True, timeout=c.get("timeout", 240))
            lines = [l for l in p.stdout.splitlines() if l.startswith("PROBE ")]
            res = json.loads(lines[-1][6:]) if lines else {"name": c["name"], "ok": False, "rc": p.returncode,
                                                           "stderr": p.stderr[-800:]}
        except subprocess.TimeoutExpired:
            res = {"name": c["name"], "ok": False, "error": "timeout"}
        res["wall_s"] = round(time.time() - t0, 1)
        line = json.dumps(res)
        print(line, flush=True)
        with open(log, "a") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
