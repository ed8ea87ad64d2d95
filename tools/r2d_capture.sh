#!/bin/bash
# 1 GPU: tests at HEAD (np>1 setup with the device-side candidate selection runs in the one-GPU multi-rank
# tests; pageable arrays take the bounce rings), bench, knobs (occupancy / line shift / unique handle), configs
O=gpurun_out; T=r2d
mkdir -p $O
python -m pytest tests -m gpu -q > $O/${T}_pytest.txt 2>&1; echo "pytest rc=$?"; tail -8 $O/${T}_pytest.txt
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"; cat $O/${T}_bench.json; tail -3 $O/${T}_bench.err
timeout 600 python tools/knob_bench.py > $O/${T}_knobs.txt 2>&1; echo "knobs rc=$?"; cat $O/${T}_knobs.txt
timeout 600 python tools/configs_bench.py > $O/${T}_configs.txt 2>&1; echo "configs rc=$?"; cat $O/${T}_configs.txt
EXAGS_SETUP_PROFILE=2 timeout 300 python tools/setup_bench_mp.py > $O/${T}_setup_mp.txt 2>&1; echo "setup mp rc=$?"; tail -20 $O/${T}_setup_mp.txt
