#!/bin/bash
# Round 2, call E: merged forward graph (per-branch continuity, pinned spin-wait, no embedding copy): tests, bench, timeline.
set +e
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -30 gpurun_out/pytest_gpu.log
echo "== bench (no extras)"
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/bench.log 2>&1
echo "bench exit $?"; tail -2 gpurun_out/bench.log | cut -c1-1500
for br in 2 4 6; do
  echo "== branches $br"
  PRIFIT_GRAPH_BRANCHES=$br timeout 300 python bench.py --steps 20 --warmup 5 --no-extras 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('branches $br: value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])"
done
echo "== graph timeline"
timeout 300 python scripts/device_timeline.py > gpurun_out/timeline_graph.txt 2>&1; echo "exit $?"; tail -6 gpurun_out/timeline_graph.txt
