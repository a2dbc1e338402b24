#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r2_pytest_gpu.log 2>&1
tail -25 gpurun_out/r2_pytest_gpu.log
python scripts/timeline.py --out gpurun_out/r2_timeline_v1.txt > /dev/null 2> gpurun_out/r2_timeline_v1.err
head -3 gpurun_out/r2_timeline_v1.txt; tail -3 gpurun_out/r2_timeline_v1.err
( time python bench.py --skip-prep --skip-wide --skip-epoch --skip-cpu-baseline ) > gpurun_out/r2_bench_v1.json 2> gpurun_out/r2_bench_v1.err
head -c 1500 gpurun_out/r2_bench_v1.json
tail -3 gpurun_out/r2_bench_v1.err
