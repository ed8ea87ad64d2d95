#!/bin/bash
# tools/run_scale.sh N [transport...]: weak-scaling bench.py on N GPUs for each transport
N=$1; shift
for t in "$@"; do
  echo "== N=$N EXAGS_EXCHANGE=$t"
  if [ "$N" = 1 ]; then
    EXAGS_EXCHANGE=$t timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu 2>gpurun_out/scale_${N}_${t}.err | tee gpurun_out/scale_${N}_${t}.json
  else
    EXAGS_EXCHANGE=$t timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) bench.py --gpus $N --steps 20 --warmup 3 2>gpurun_out/scale_${N}_${t}.err | tee gpurun_out/scale_${N}_${t}.json
  fi
  tail -3 gpurun_out/scale_${N}_${t}.err
done
