#!/bin/bash
# Round 2, call B: tests after the fit rewrite + speculative backward, bench, graph timeline, rows-cluster experiment.
set +e
mkdir -p gpurun_out
PYTEST_TAIL=40 BENCH_ARGS="--steps 20 --warmup 5" bash scripts/gpu_round.sh
echo "== graph timeline"
timeout 300 python scripts/device_timeline.py > gpurun_out/timeline_graph.txt 2>&1; echo "exit $?"; tail -8 gpurun_out/timeline_graph.txt
for c in 6 8; do
  echo "== PRIFIT_ROWS_CLUSTER=$c"
  PRIFIT_ROWS_CLUSTER=$c timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_cluster$c.log 2>&1; echo "exit $?"
  python - <<PY
import json
for l in open("gpurun_out/bench_cluster$c.log"):
    if l.startswith("{"):
        d = json.loads(l); print("cluster $c: value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "eager", d["config"]["eager_ms_per_step"])
PY
done
echo "== ncu full: fit kernels after the rewrite"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fit_fwd|fit_bwd" -s 6 -c 2 -f \
    -o gpurun_out/prof_r02_fit_after python bench.py --steps 2 --warmup 3 --no-graph > gpurun_out/ncu_fit_after.log 2>&1
echo "ncu exit $?"
