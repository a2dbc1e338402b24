#!/bin/bash
mkdir -p gpurun_out
for cfg in 8,16,8,8 16,32,16,16; do
  tag=$(echo $cfg | tr , _)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_head_launches_$tag.csv python scripts/decode_probe.py $cfg 1024 1 > gpurun_out/r2_head_launches_$tag.log 2>&1
  python scripts/agg_launches.py gpurun_out/r2_head_launches_$tag.csv 2>/dev/null | head -14
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_decode_head -c 1 -o gpurun_out/r2_head_A python scripts/decode_probe.py 8,16,8,8 1024 1 > gpurun_out/r2_head_ncuA.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_decode_head -c 1 -o gpurun_out/r2_head_B python scripts/decode_probe.py 16,32,16,16 1024 1 > gpurun_out/r2_head_ncuB.log 2>&1
ls -la gpurun_out/*.ncu-rep
