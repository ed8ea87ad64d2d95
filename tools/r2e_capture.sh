#!/bin/bash
# 8 GPUs: multi-GPU parity at HEAD (every np <= 8, both transports), config 5 (strong scaling, 800 M DOF N=9)
# with all 16 dom x op checks, its stage times, and the weak-scaling bench with the check and NVLink counters
O=gpurun_out; T=r2e
mkdir -p $O
python -m pytest tests/test_multi_gpu.py -q -k "not one_gpu" > $O/${T}_pytest_8gpu.txt 2>&1; echo "pytest rc=$?"; tail -6 $O/${T}_pytest_8gpu.txt
bash tools/run_strong.sh 8 ${T}_P8 --steps 10
EXAGS_TRACE=1 bash tools/run_strong.sh 8 ${T}_P8_trace --steps 10 --only add,double --no-check | grep -h "exags trace" | head -8 > $O/${T}_strong_trace.txt; cat $O/${T}_strong_trace.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 > $O/${T}_bench_N8.json 2> $O/${T}_bench_N8.err; echo "bench rc=$?"; cat $O/${T}_bench_N8.json; tail -3 $O/${T}_bench_N8.err
