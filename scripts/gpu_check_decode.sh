#!/bin/bash
# quick check after a decode-path change: decode / codec / step parity tests, device-resident decode times of the two
# configurations, per-kernel times of one decode each
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py tests/test_gpu_codec.py tests/test_gpu_step.py tests/test_gpu_fit.py -x -q ) > gpurun_out/r2_head_pytest.log 2>&1
tail -3 gpurun_out/r2_head_pytest.log
timeout 300 python scripts/decode_probe.py 8,16,8,8 1024 2>&1 | tail -1
timeout 300 python scripts/decode_probe.py 16,32,16,16 1024 2>&1 | tail -1
bash scripts/gpu_decode_launches.sh
