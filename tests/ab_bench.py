#!/usr/bin/env python3
"""Interleaved A/B timing of kernel variants on ONE box (clock-normalised).

    python tests/ab_bench.py B N k rounds name=ENV1=v,ENV2=v name2= ...

Variants are run round-robin; each call is timed with CUDA events while a thread samples the SM
clock through NVML.  Prints per-variant median ms, median SM MHz and ms normalised to 1350 MHz.
"""
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "crowd-coachable-recommendations_b200"))
import pynvml  # noqa: E402
import torch  # noqa: E402

from ccr_b200 import engine  # noqa: E402

B, N, k, rounds = (int(x) for x in sys.argv[1:5])
variants = []
for spec in sys.argv[5:]:
    name, _, envs = spec.partition("=")
    env = dict(e.split("=") for e in envs.split(",") if e) if envs else {}
    variants.append((name, env))
algo = int(os.environ.get("AB_ALGO", "2"))
mask_h = int(os.environ.get("AB_MASK", "0"))
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
items = torch.empty((N, 768), dtype=torch.bfloat16, device=dev)
for s in range(0, N, 1 << 20):
    e = min(N, s + (1 << 20))
    items[s:e] = torch.randn((e - s, 768), generator=g, device=dev).to(torch.bfloat16)
q = torch.randn((B, 768), generator=g, device=dev).to(torch.bfloat16)
mask = None
if mask_h:
    import numpy as np
    rs = np.random.RandomState(2)
    if mask_h < 0:
        rows = [np.zeros(0, dtype=np.int64) for _ in range(B)]
    else:
        rows = [np.unique(rs.randint(0, N, size=min(64, rs.geometric(1.0 / mask_h)))) for _ in range(B)]
    mask = engine.SparseMask.from_lists(rows, N, -1e6, engine.MASK_SET, dev)

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples, stop = [], threading.Event()


def sampler():
    while not stop.is_set():
        samples.append((time.time(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                        pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
        time.sleep(0.005)


th = threading.Thread(target=sampler, daemon=True)
th.start()
ALL_KEYS = {k_ for _, env in variants for k_ in env}
res = {name: [] for name, _ in variants}
for r in range(rounds + 1):
    for name, env in variants:
        for k_ in ALL_KEYS:
            os.environ.pop(k_, None)
        os.environ.update(env)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.time()
        ev0.record()
        engine.score_topk(q, items, k, mask=mask, algo=algo)
        ev1.record()
        torch.cuda.synchronize()
        t1 = time.time()
        clk = [c for (t, c, p) in samples if t0 <= t <= t1]
        if r > 0:  # round 0 = warm-up
            res[name].append((ev0.elapsed_time(ev1), statistics.median(clk) if clk else float("nan")))
stop.set()
for name, _ in variants:
    ms = statistics.median(x[0] for x in res[name])
    mhz = statistics.median(x[1] for x in res[name])
    print(f"AB {name:12s} B={B} N={N} k={k}: median {ms:8.3f} ms  @ {mhz:6.0f} MHz  -> {ms * mhz / 1350:8.3f} ms@1350  "
          f"qps {B / ms * 1e3:9.0f}  min {min(x[0] for x in res[name]):.3f}")
