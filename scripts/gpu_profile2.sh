#!/bin/bash
# ncu --set full captures inside the eager bench (one process per capture):  CAPS="name:regex:skip ..." bash scripts/gpu_profile2.sh
set +e
mkdir -p gpurun_out
for cap in ${CAPS}; do
  name=${cap%%:*}; rest=${cap#*:}; regex=${rest%%:*}; skip=${rest#*:}
  echo "== ncu full capture ${name}: kernel regex ${regex}, skip ${skip}"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:${regex} -s ${skip} -c 1 -f \
      -o gpurun_out/prof_${name} python bench.py --steps 2 --warmup 3 --no-graph ${BENCH_ARGS} > gpurun_out/ncu_full_${name}.log 2>&1
  echo "ncu exit $?"
done
