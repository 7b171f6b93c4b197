#!/bin/bash
# usage: tests/ab/clk_run.sh <label> <env...> -- args to ncu_case.py ; samples SM clock / power while running
label=$1; shift
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.active --format=csv,noheader -lms 100 > /tmp/clk_$label.csv &
SMI=$!
env "$@"
kill $SMI
python - <<PY
import re
rows=[l.strip().split(", ") for l in open("/tmp/clk_$label.csv") if l.strip()]
clk=[int(r[0].split()[0]) for r in rows]; pw=[float(r[1].split()[0]) for r in rows]
hi=[c for c,p in zip(clk,pw) if p>600]
print("$label", "samples",len(clk),"under-load",len(hi),"clk median under load", sorted(hi)[len(hi)//2] if hi else None, "min", min(hi) if hi else None, "power max", max(pw), "reasons", set(r[2] for r in rows))
PY
