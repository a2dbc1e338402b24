#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/r2_pytest_gpu.log 2>&1
tail -6 gpurun_out/r2_pytest_gpu.log
NVF_ROWS=0 python scripts/timeline.py --out gpurun_out/r2_timeline_rows0.txt > /dev/null 2> gpurun_out/r2_timeline_rows0.err
python scripts/timeline.py --out gpurun_out/r2_timeline_rows1.txt > /dev/null 2> gpurun_out/r2_timeline_rows1.err
grep -E "step wall|convT5_fwd" gpurun_out/r2_timeline_rows0.txt gpurun_out/r2_timeline_rows1.txt
python bench.py --skip-prep --skip-wide --skip-epoch --skip-cpu-baseline 2> gpurun_out/r2_bench_c.err | head -c 700
