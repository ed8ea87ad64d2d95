#!/bin/bash
# third GPU call of round 2 (2 GPUs): multi-GPU parity at HEAD (2-rank cases), bench at N=2 with the
# check, latency table at np=2, ncu of the np>1 kernels with NVLink counters
O=gpurun_out; T=r2c
mkdir -p $O
python -m pytest tests/test_multi_gpu.py -q -k "not one_gpu" > $O/${T}_pytest_2gpu.txt 2>&1; echo "pytest rc=$?"; tail -5 $O/${T}_pytest_2gpu.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > $O/${T}_bench_N2.json 2> $O/${T}_bench_N2.err; echo "bench rc=$?"; cat $O/${T}_bench_N2.json; tail -3 $O/${T}_bench_N2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/latency_table.py > $O/${T}_latency_np2.txt 2>&1; echo "latency rc=$?"; cat $O/${T}_latency_np2.txt
bash tools/ncu_np2.sh $O/${T}_np2
EXAGS_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 10 > $O/${T}_trace_N2.txt 2>&1; grep -h "exags trace" $O/${T}_trace_N2.txt
