#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_step.py tests/test_gpu_fit.py tests/test_gpu_grids.py -x -q ) > gpurun_out/r2_train_pytest.log 2>&1
tail -3 gpurun_out/r2_train_pytest.log
for v in 0 1; do
  NVF_TRAIN_POLY=$v python bench.py --skip-prep --skip-wide --skip-epoch --skip-cpu-baseline --steps 20 2> gpurun_out/r2_bench_poly$v.err | head -c 330; echo
done
NVF_TRAIN_POLY=1 python scripts/timeline.py --out gpurun_out/r2_timeline_poly1.txt > /dev/null 2> gpurun_out/r2_timeline_poly1.err
grep -E "poly|pad_in|convT5_fwd" gpurun_out/r2_timeline_poly1.txt | cut -c1-120
