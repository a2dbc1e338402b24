#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r2_pytest_gpu.log 2>&1
tail -4 gpurun_out/r2_pytest_gpu.log
python scripts/timeline.py --out gpurun_out/r2_timeline_v2.txt > /dev/null 2> gpurun_out/r2_timeline_v2.err
head -1 gpurun_out/r2_timeline_v2.txt
( time python bench.py --skip-prep --skip-wide --skip-epoch --skip-cpu-baseline ) 2> gpurun_out/r2_bench_e.err | head -c 900
echo
( time python bench.py --skip-prep --skip-wide --skip-epoch --skip-cpu-baseline --flush-l2 ) 2>> gpurun_out/r2_bench_e.err | head -c 500
