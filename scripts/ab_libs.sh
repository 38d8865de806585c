#!/bin/bash
# A/B of two builds of the library on ONE box, alternating: prifit_b200/libprifit_b200_prev.so against libprifit_b200_new.so
# (both git-ignored; the one left in place at the end is _new).  Prints shapes/s resident, ms per step, shapes/s end to end.
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for round in 1; do
  for which in new prev; do
    cp prifit_b200/libprifit_b200_${which}.so prifit_b200/libprifit_b200.so
    if [ $round = 1 ]; then echo "== ${which}: gram kernels"; python scripts/time_gram.py 2>&1 | grep -v Warn | tail -6; fi
    python bench.py --steps 200 --warmup 5 --no-extras --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('${which}', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'eager', d['config'].get('eager_ms_per_step'))"
  done
done
cp prifit_b200/libprifit_b200_new.so prifit_b200/libprifit_b200.so
