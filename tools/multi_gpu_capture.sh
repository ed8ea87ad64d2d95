#!/bin/bash
# tools/multi_gpu_capture.sh TAG N [strong] -- the multi-GPU evidence of a round in one gpurun call:
#   gpurun --gpus 8 --timeout 1500 -- 'bash tools/multi_gpu_capture.sh r2j 8 strong'
# 1. tests/test_multi_gpu.py without the one-GPU variants (every np <= N, both transports) -> TAG_pytest_Ngpu.txt
# 2. bench.py under torchrun on N GPUs (weak scaling, with the check)                      -> TAG_bench_NN.json
# 3. "strong": config 5 (800 M DOF N=9, irregular blocks) on N GPUs, all 16 dom x op cases -> strong_TAG_PN.json,
#    and its per-stage times with the stages serialised (EXAGS_TRACE=1)                    -> TAG_strong_trace.txt
# NVLink bytes of the send kernel: tools/nvlink_pack_probe (2 GPUs, one process) under ncu, see its header.
T=${1:-rX}; N=${2:-2}; O=gpurun_out
mkdir -p $O
python -m pytest tests/test_multi_gpu.py -q -k "not one_gpu" > $O/${T}_pytest_${N}gpu.txt 2>&1; echo "pytest rc=$?"; tail -6 $O/${T}_pytest_${N}gpu.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N \
    > $O/${T}_bench_N${N}.json 2> $O/${T}_bench_N${N}.err; echo "bench rc=$?"; cut -c1-700 $O/${T}_bench_N${N}.json; tail -3 $O/${T}_bench_N${N}.err
if [ "$3" = "strong" ]; then
  bash tools/run_strong.sh $N ${T}_P${N} --steps 10
  EXAGS_TRACE=1 bash tools/run_strong.sh $N ${T}_P${N}_trace --steps 10 --only add,double --no-check | tr '[' '\n' | grep "exags trace" | sed 's/^/[/' > $O/${T}_strong_trace.txt
  cat $O/${T}_strong_trace.txt
fi
