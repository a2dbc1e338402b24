#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_step.py -m gpu -q -x 2>&1 | tail -3
for cfg in "2 4" "4 8" "1 2"; do
  set -- $cfg
  NVF_KS_UP2=$1 NVF_KS_UP1=$2 python scripts/timeline.py --out gpurun_out/r2_tl_ks$1_$2.txt > /dev/null 2> gpurun_out/r2_tl.err
  echo "KS_UP2=$1 KS_UP1=$2"; grep -E "step wall|convT5_fwd" gpurun_out/r2_tl_ks$1_$2.txt
done
