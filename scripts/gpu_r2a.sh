#!/bin/bash
# Round 2, call A: GPU parity tests (new fixtures), smoke, bench (both arms), ncu --set full of the reduction kernels, topology.
set +e
mkdir -p gpurun_out
PYTEST_TAIL=80 BENCH_ARGS="--steps 20 --warmup 5" bash scripts/gpu_round.sh
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
echo "ref exit $?"; tail -2 gpurun_out/bench_ref.log
echo "== ncu full: reduction kernels (before the round-2 rebuild)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fit_fwd|fit_bwd|sdf_fwd|sdf_bwd|sdf_finalize|membership_" -s 30 -c 9 -f \
    -o gpurun_out/prof_r02_reductions_before python bench.py --steps 2 --warmup 3 --no-graph > gpurun_out/ncu_red_before.log 2>&1
echo "ncu exit $?"
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
(lscpu | head -30; numactl -H 2>&1; nproc) >> gpurun_out/topo.txt 2>&1
