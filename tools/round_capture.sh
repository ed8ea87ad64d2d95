#!/bin/bash
# tools/round_capture.sh TAG -- the standard evidence of one round in ONE gpurun call (1 GPU):
#   gpurun --timeout 1500 -- 'bash tools/round_capture.sh r2i'
# 1. pytest -m gpu, smoke()   -> gpurun_out/TAG_pytest.txt, TAG_smoke.txt
# 2. bench.py (default flags) and its reference arm -> TAG_bench.json, TAG_bench_ref.json (never under a profiler)
# 3. every other configuration -> TAG_configs.txt (tools/configs_bench.py); layout knobs -> TAG_knobs.txt
# 4. one ncu --set full launch of every kernel on that list -> TAG_configs.ncu-rep (read here with
#    tools/ncu_summary.py), and the launch list of the bench command -> TAG_launches.csv
# Copy what is to be judged to profiles/.
T=${1:-rX}; O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q > $O/${T}_pytest.txt 2>&1; echo "pytest rc=$?"; tail -6 $O/${T}_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $O/${T}_smoke.txt
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; rc=$?; echo "bench rc=$rc"; cut -c1-600 $O/${T}_bench.json; tail -3 $O/${T}_bench.err
python bench.py --impl reference --steps 20 --warmup 3 > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err; echo "ref rc=$?"; cut -c1-400 $O/${T}_bench_ref.json
timeout 600 python tools/configs_bench.py > $O/${T}_configs.txt 2>&1; echo "configs rc=$?"; cat $O/${T}_configs.txt
if [ "$2" = "ncu" ] && [ $rc = 0 ]; then
  timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o $O/${T}_configs \
      python tools/configs_bench.py --ncu > $O/${T}_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 $O/${T}_ncu.log
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu > $O/${T}_ncu_launches.log 2>&1; echo "launch list rc=$?"
fi
