#!/bin/bash
mkdir -p gpurun_out
NVF_POLY_A=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_convT5_poly -s 2 -c 1 -o gpurun_out/r2_polyA python scripts/decode_probe.py 8,16,8,8 1024 1 > gpurun_out/r2_poly_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_convT5_fwd -s 2 -c 1 -o gpurun_out/r2_tileA python scripts/decode_probe.py 8,16,8,8 1024 1 >> gpurun_out/r2_poly_ncu.log 2>&1
ls -la gpurun_out/r2_polyA.ncu-rep gpurun_out/r2_tileA.ncu-rep
