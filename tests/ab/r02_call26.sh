#!/bin/bash
# round 2, GPU call 26: last build -- full GPU suite + smoke + default bench.py
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q --maxfail=10 ) > $O/r02_c26_pytest.log 2>&1; echo "pytest rc $?" >> $O/r02_c26_pytest.log
tail -7 $O/r02_c26_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_c26_smoke.log 2>&1; tail -3 $O/r02_c26_smoke.log
( time timeout 900 python bench.py ) > $O/r02_c26_bench.json 2> $O/r02_c26_bench.err; cut -c 1-400 $O/r02_c26_bench.json; tail -4 $O/r02_c26_bench.err
