#!/bin/bash
# tools/round_capture.sh TAG -- the standard evidence of one round in ONE gpurun call (1 GPU):
#   gpurun --timeout 900 -- 'bash tools/round_capture.sh r2a'
# 1. pytest -m gpu            -> gpurun_out/pytest_TAG.txt
# 2. bench.py (default flags) -> gpurun_out/bench_TAG.json        (never under a profiler)
# 3. the other configs        -> gpurun_out/configs_TAG.txt, weighted_TAG.txt
# 4. ncu launch list of the bench command (only after 2 exited 0) -> launches_TAG.csv
# 5. one `ncu --set full` capture of the hot kernel               -> prof_TAG.ncu-rep
# Read the .ncu-rep on the CPU box with tools/ncu_summary.py and copy what is to be judged to profiles/.
TAG=${1:-rX}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_$TAG.txt
python bench.py --steps 20 --warmup 3 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; rc=$?; echo "bench rc=$rc"; cat $O/bench_$TAG.json
timeout 300 python tools/configs_bench.py > $O/configs_$TAG.txt 2>&1; cat $O/configs_$TAG.txt
timeout 120 python tools/weighted_bench.py > $O/weighted_$TAG.txt 2>&1; cat $O/weighted_$TAG.txt
if [ $rc = 0 ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu > $O/ncu_${TAG}_a.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_fused_sell -s 3 -c 1 -o $O/prof_$TAG -f \
      python bench.py --steps 2 --warmup 3 --no-cpu > $O/ncu_${TAG}_b.log 2>&1
  tail -2 $O/ncu_${TAG}_b.log
fi
