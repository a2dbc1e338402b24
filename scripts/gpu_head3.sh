#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py tests/test_gpu_codec.py -x -q -k "decode or encdec or codec or cloud" ) > gpurun_out/r2_head_pytest.log 2>&1
tail -3 gpurun_out/r2_head_pytest.log
NVF_DECODE_HEAD=1 timeout 300 python scripts/decode_probe.py 8,16,8,8 1024 2>&1 | tail -1
NVF_DECODE_HEAD=1 timeout 300 python scripts/decode_probe.py 16,32,16,16 1024 2>&1 | tail -1
for cfg in 8,16,8,8 16,32,16,16; do
  tag=$(echo $cfg | tr , _)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_decode_head -s 2 -c 1 -o gpurun_out/r2_head_$tag python scripts/decode_probe.py $cfg 1024 1 > gpurun_out/r2_head_ncu_$tag.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
