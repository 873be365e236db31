#!/bin/bash
# one strong-scaling point of bench.py, launched the way the driver launches it:
#   gpurun --gpus N --timeout 600 -- 'bash scripts/scale_bench.sh N [extra bench flags]'
n=$1; shift
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $n --steps 50 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
echo "rc=$?"; tail -3 gpurun_out/scale_n$n.err; cut -c1-900 gpurun_out/scale_n$n.json
