#!/bin/bash
# 1 GPU, HEAD: the whole gpu test tier, bench, all configs, one ncu --set full launch of every kernel
O=gpurun_out; T=r2h
mkdir -p $O
python -m pytest tests -m gpu -q > $O/${T}_pytest.txt 2>&1; echo "pytest rc=$?"; tail -6 $O/${T}_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $O/${T}_smoke.txt
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"; cat $O/${T}_bench.json | cut -c1-600; tail -3 $O/${T}_bench.err
python bench.py --impl reference --steps 10 --warmup 2 > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err; echo "ref rc=$?"; cat $O/${T}_bench_ref.json | cut -c1-400
timeout 600 python tools/configs_bench.py > $O/${T}_configs.txt 2>&1; echo "configs rc=$?"; cat $O/${T}_configs.txt
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o $O/${T}_configs \
    python tools/configs_bench.py --ncu > $O/${T}_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 $O/${T}_ncu.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/${T}_ncu_launches.log 2>&1; echo "launch list rc=$?"
