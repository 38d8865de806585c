#!/bin/bash
# ncu --set full captures of named kernels inside bench.py (one process per kernel regex).
#   KERNELS="meanshift_tc_kernel fit_fwd_kernel" bash scripts/gpu_profile.sh
# Reports land in gpurun_out/prof_<kernel>.ncu-rep; read here with `ncu -i ... --page raw --csv`.
set +e
mkdir -p gpurun_out
for k in ${KERNELS}; do
  echo "== ncu full capture of ${k}"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:${k} -s ${NCU_SKIP:-4} -c 1 -f \
      -o gpurun_out/prof_${k} python bench.py --steps 2 --warmup 3 ${BENCH_ARGS} > gpurun_out/ncu_full_${k}.log 2>&1
  echo "ncu full exit $?"
done
