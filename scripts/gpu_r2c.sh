#!/bin/bash
# Round 2, call C: log-binned bandwidth (2 Gram passes), intersection kernels; tests, bench, Gram kernel times.
set +e
mkdir -p gpurun_out
PYTEST_TAIL=40 BENCH_ARGS="--steps 20 --warmup 5" bash scripts/gpu_round.sh
echo "== gram kernel times cfg2"
timeout 300 python scripts/time_gram.py 2>&1 | grep -v Warn | tail -8
echo "== gram kernel times cfg4"
B=16 N=10000 timeout 300 python scripts/time_gram.py 2>&1 | grep -v Warn | tail -8
echo "== graph timeline"
timeout 300 python scripts/device_timeline.py > gpurun_out/timeline_graph.txt 2>&1; echo "exit $?"; tail -6 gpurun_out/timeline_graph.txt
