#!/bin/bash
# Evidence of the final state of a round, one gpurun call: GPU parity tests, smoke, the default bench line, the ncu launch
# list of the same command, the device timeline of the graph step.  Logs -> gpurun_out/ (copied into profiles/ by hand).
set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?"; grep "smoke engine" gpurun_out/smoke.log
echo "== bench (default line)"
timeout 900 python bench.py > gpurun_out/bench_full.log 2>&1
echo "bench exit $?"; tail -1 gpurun_out/bench_full.log | cut -c1-400
echo "== ncu launch list (eager steps + graph replays of the same command)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-graph > gpurun_out/ncu_bench.log 2>&1
echo "ncu exit $?"
echo "== graph timeline"
timeout 300 python scripts/device_timeline.py > gpurun_out/timeline_graph.txt 2>&1; echo "exit $?"; tail -5 gpurun_out/timeline_graph.txt
