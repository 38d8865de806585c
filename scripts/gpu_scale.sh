#!/bin/bash
# bench.py under torchrun on N GPUs (run with gpurun --gpus N):  NGPU=2 bash scripts/gpu_scale.sh
set +e
mkdir -p gpurun_out
N=${NGPU:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.log 2>&1
echo "bench n=$N exit $?"; tail -2 gpurun_out/bench_n$N.log
