#!/bin/bash
# round 2, GPU call 2: full GPU suite, throttle-lead / CTA-variant sweep, bench.py, crash repro + memcheck
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > $O/r02_c2_pytest.log 2>&1; echo "pytest rc $?" >> $O/r02_c2_pytest.log
tail -5 $O/r02_c2_pytest.log
timeout 900 python tests/perf_sweep.py --batches 256,384,512,640,768,1024,1536,2048,3072,4096,8192,16384 \
  --variants "default=;lead2=CCR_LEAD=2;lead4=CCR_LEAD=4;lead8=CCR_LEAD=8;lead16=CCR_LEAD=16;thr1=CCR_THROTTLE=1;nothr=CCR_THROTTLE=0;pair=CCR_2CTA=1;single=CCR_2CTA=0" \
  --secs 0.3 --rounds 2 --md $O/r02_c2_sweep_lead.md > $O/r02_c2_sweep_lead.log 2>&1; tail -120 $O/r02_c2_sweep_lead.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r02_c2_bench.json 2> $O/r02_c2_bench.err; tail -c 3000 $O/r02_c2_bench.json; tail -3 $O/r02_c2_bench.err
echo "== crash repro (seeded default)"; timeout 120 python tests/crash_case.py 260 280000 768 1001 2 2>&1 | tail -4
echo "== crash repro (CCR_NO_SEED=1)"; CCR_NO_SEED=1 timeout 120 python tests/crash_case.py 260 280000 768 1001 2 2>&1 | tail -4
echo "== memcheck"; CCR_NO_SEED=1 timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tests/crash_case.py 260 280000 768 1001 2 > $O/r02_c2_memcheck.log 2>&1; grep -m 40 -E "Invalid|at |by thread|Address|ERROR SUMMARY|ok " $O/r02_c2_memcheck.log
