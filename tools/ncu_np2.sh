#!/bin/bash
# tools/ncu_np2.sh OUT  -- ncu --set full + NVLink byte counters on rank 0 of a 2-rank gs call
OUT=${1:-gpurun_out/np2}
export MASTER_ADDR=127.0.0.1 MASTER_PORT=29571 WORLD_SIZE=2 EXAGS_JOB_KEY=ncunp2$$ EXAGS_PUSH_TIMEOUT_S=900
RANK=1 LOCAL_RANK=1 python tools/np2_step.py > ${OUT}_rank1.log 2>&1 &
P1=$!
RANK=0 LOCAL_RANK=0 timeout 900 ncu --profile-from-start off --set full \
    --metrics nvltx__bytes.sum,nvlrx__bytes.sum,nvltx__bytes_data_user.sum,nvlrx__bytes_data_user.sum \
    --clock-control none --import-source on -f -o ${OUT} python tools/np2_step.py > ${OUT}_rank0.log 2>&1
echo "ncu rank0 rc=$?"; tail -4 ${OUT}_rank0.log
wait $P1; echo "rank1 rc=$?"
