#!/bin/bash
O=gpurun_out; T=r2f
python -m pytest tests -m gpu -q -x -k "weighted or flagged or unique or all_domains or golden or known or execw or strong" > $O/${T}_pytest.txt 2>&1; echo "pytest rc=$?"; tail -6 $O/${T}_pytest.txt
timeout 300 python tools/configs_bench.py --only "unique" > $O/${T}_configs_unique.txt 2>&1; cat $O/${T}_configs_unique.txt
timeout 300 python tools/configs_bench.py --only "1-in-7" > $O/${T}_configs_flagged.txt 2>&1; cat $O/${T}_configs_flagged.txt
bash tools/run_strong.sh 1 ${T}_P1 --steps 10 --only add,double
