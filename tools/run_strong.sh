#!/bin/bash
# tools/run_strong.sh P tag [strong_bench args...]: config-5 strong-scaling case on P GPUs
P=$1; tag=$2; shift 2
if [ "$P" = 1 ]; then
  timeout 1500 python tools/strong_bench.py --out gpurun_out/strong_${tag}.json "$@" 2>gpurun_out/strong_${tag}.err | tail -20
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) tools/strong_bench.py --out gpurun_out/strong_${tag}.json "$@" 2>gpurun_out/strong_${tag}.err | tail -20
fi
tail -3 gpurun_out/strong_${tag}.err
