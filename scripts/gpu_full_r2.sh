#!/bin/bash
# full GPU tier + default bench lines (native and reference arm) as the driver runs them
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r2_pytest_gpu.log 2>&1
tail -4 gpurun_out/r2_pytest_gpu.log
( time python bench.py ) > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
tail -c 1500 gpurun_out/r2_bench_default.json; echo
tail -3 gpurun_out/r2_bench_default.err
