#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/decode_probe.py 8,16,8,8 1024 2>&1 | tail -1
timeout 300 python scripts/decode_probe.py 16,32,16,16 1024 2>&1 | tail -1
bash scripts/gpu_head5.sh 2>&1 | grep -E "head|poly|last"
