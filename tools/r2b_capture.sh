#!/bin/bash
# second GPU call of round 2 (1 GPU): tests (incl. the persistent TMA kernels, np>1 CUDA graph on one GPU,
# comm_dot_async), bench with the pageable-host legs, configs, knobs (pipe / line shift), latency table,
# ncu of every kernel on the list with the pipelined form
O=gpurun_out; T=r2b
mkdir -p $O
python -m pytest tests -m gpu -q > $O/${T}_pytest.txt 2>&1; echo "pytest rc=$?"; tail -8 $O/${T}_pytest.txt
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"; cat $O/${T}_bench.json; tail -3 $O/${T}_bench.err
timeout 600 python tools/knob_bench.py > $O/${T}_knobs.txt 2>&1; echo "knobs rc=$?"; cat $O/${T}_knobs.txt
timeout 600 python tools/configs_bench.py > $O/${T}_configs.txt 2>&1; echo "configs rc=$?"; cat $O/${T}_configs.txt
EXAGS_SELL_PIPE=2 timeout 600 python tools/configs_bench.py > $O/${T}_configs_pipe2.txt 2>&1; echo "configs pipe2 rc=$?"; cat $O/${T}_configs_pipe2.txt
timeout 600 python tools/latency_table.py > $O/${T}_latency.txt 2>&1; echo "latency rc=$?"; cat $O/${T}_latency.txt
EXAGS_SELL_PIPE=2 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o $O/${T}_configs_pipe2 \
    python tools/configs_bench.py --ncu > $O/${T}_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $O/${T}_ncu.log
