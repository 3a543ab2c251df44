#!/bin/bash
# ncu evidence of one round-2 step: launch list (+DRAM bytes) and --set full of every conv launch.
mkdir -p gpurun_out
K='regex:conv_|decode|nms|im2col|pack|maxpool|spp|plan_dest|emit'
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" \
  --csv --log-file gpurun_out/${TAG}_launches.csv python tools/profile_step.py --steps 1 > gpurun_out/${TAG}_launches.log 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/${TAG}_launches.log
timeout 1500 ncu --set full --clock-control none -k 'regex:conv_' --launch-skip 73 --launch-count 73 \
  -f -o /tmp/${TAG}_conv_full python tools/profile_step.py --steps 1 > gpurun_out/${TAG}_conv_full.log 2>&1
echo "conv full rc=$?"; tail -2 gpurun_out/${TAG}_conv_full.log
# the report itself is > 64 MiB: keep the raw page as CSV (what tools/summarize_profiles.py reads)
ncu -i /tmp/${TAG}_conv_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_conv_full_raw.csv 2> gpurun_out/${TAG}_conv_full_raw.err
ls -la /tmp/${TAG}_conv_full.ncu-rep gpurun_out/${TAG}_conv_full_raw.csv
# memory-bound kernels of the other configs: spp (608, B=32) and tiny (max-pools, u8 packing), dense decode
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k 'regex:spp|maxpool|im2col|pack|decode_dense' --csv --log-file gpurun_out/${TAG}_hbm_spp608.csv \
  python tools/profile_step.py --cfg yolov3-spp --size 608 --batch 32 --steps 1 --dense > gpurun_out/${TAG}_hbm_spp608.log 2>&1
echo "spp rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k 'regex:spp|maxpool|im2col|pack|decode_dense' --csv --log-file gpurun_out/${TAG}_hbm_tiny416.csv \
  python tools/profile_step.py --cfg yolov3-tiny --size 416 --batch 64 --steps 1 --dense > gpurun_out/${TAG}_hbm_tiny416.log 2>&1
echo "tiny rc=$?"
