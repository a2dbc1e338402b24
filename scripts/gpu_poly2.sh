#!/bin/bash
mkdir -p gpurun_out
for cfg in 8,16,8,8 16,32,16,16; do
  tag=$(echo $cfg | tr , _)
  NVF_POLY_A=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_poly_$tag.csv python scripts/decode_probe.py $cfg 1024 1 2>&1 | tail -1
  grep -E "poly|convT5_fwd" gpurun_out/r2_poly_$tag.csv | tail -1 | awk -F'","' '{print $5, $(NF)}'
done
