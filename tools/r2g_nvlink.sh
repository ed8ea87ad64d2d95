#!/bin/bash
O=gpurun_out
for ns in 201601 4363087; do
  tools/nvlink_pack_probe $ns
  timeout 300 ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,nvltx__bytes_data_user.sum,nvlrx__bytes_data_user.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum \
      --clock-control none -k regex:k_pack_slots -s 3 -c 2 --csv --log-file $O/r2g_nvlink_pack_${ns}.csv tools/nvlink_pack_probe $ns > $O/r2g_nvlink_pack_${ns}.log 2>&1
  echo "ncu rc=$?"; tail -12 $O/r2g_nvlink_pack_${ns}.csv
done
