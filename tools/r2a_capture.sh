#!/bin/bash
# first GPU call of round 2 (1 GPU): full gpu test tier, bench with the new check, all configs,
O=gpurun_out; T=r2a
mkdir -p $O
python -m pytest tests -m gpu -q > $O/${T}_pytest.txt 2>&1; echo "pytest rc=$?"; tail -5 $O/${T}_pytest.txt
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"; cat $O/${T}_bench.json; tail -3 $O/${T}_bench.err
timeout 600 python tools/configs_bench.py > $O/${T}_configs.txt 2>&1; echo "configs rc=$?"; cat $O/${T}_configs.txt
timeout 600 python tools/knob_bench.py > $O/${T}_knobs.txt 2>&1; echo "knobs rc=$?"; cat $O/${T}_knobs.txt
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o $O/${T}_configs \
    python tools/configs_bench.py --ncu > $O/${T}_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $O/${T}_ncu.log
ls -la $O | tail -12
