#!/bin/bash
# One gpurun call: tcgen05 probe scan, GPU parity tests, smoke, a short bench and an ncu launch list.
# Logs -> gpurun_out/.  Every step runs under its own timeout.
set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
if [ -n "${PROBE}" ]; then
  echo "== tcgen05 probe scan"
  timeout 900 python scripts/probe_scan.py
fi
echo "== pytest -m gpu ${PYTEST_ARGS}"
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -${PYTEST_TAIL:-40} gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?"; grep -v Warning gpurun_out/smoke.log | tail -5
echo "== bench ${BENCH_ARGS}"
timeout 600 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.log 2>&1
echo "bench exit $?"; tail -3 gpurun_out/bench.log
if [ -n "${NCU_LIST}" ]; then
  echo "== ncu launch list"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 ${BENCH_ARGS} > gpurun_out/ncu_bench.log 2>&1
  echo "ncu exit $?"
fi
if [ -n "${NCU_FULL}" ]; then
  echo "== ncu full capture of ${NCU_FULL}"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:${NCU_FULL} -s 3 -c 2 -f -o gpurun_out/prof_${NCU_FULL} \
      python bench.py --steps 2 --warmup 3 ${BENCH_ARGS} > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?"
fi
