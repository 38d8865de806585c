#!/bin/bash
# Round 2, call D: cfg5 tests + full default bench line (sustained, hbm, cfg4, cfg5 sub-benches), Gram timings after the ALU trim.
set +e
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -40 gpurun_out/pytest_gpu.log
echo "== default bench"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1
echo "bench exit $?"; tail -3 gpurun_out/bench.log
echo "== gram kernel times cfg2"
timeout 300 python scripts/time_gram.py 2>&1 | grep -v Warn | tail -6
