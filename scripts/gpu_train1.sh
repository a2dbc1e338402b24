#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_step.py tests/test_gpu_fit.py -x -q ) > gpurun_out/r2_train_pytest.log 2>&1
tail -3 gpurun_out/r2_train_pytest.log
for hi in 0 1; do
  NVF_HI=$hi python bench.py --skip-prep --skip-wide --skip-epoch --skip-cpu-baseline --steps 20 2> gpurun_out/r2_bench_hi$hi.err | head -c 400; echo
done
NVF_HI=1 python scripts/timeline.py --out gpurun_out/r2_timeline_hi1.txt > /dev/null 2> gpurun_out/r2_timeline_hi1.err
head -1 gpurun_out/r2_timeline_hi1.txt
