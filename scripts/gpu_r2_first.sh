#!/bin/bash
# round 2, first GPU call: parity tier, bench line, concurrent timeline, half-batch probe
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_smi.txt 2>&1
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest_gpu.log 2>&1
tail -5 gpurun_out/r2_pytest_gpu.log
python scripts/timeline.py --out gpurun_out/r2_timeline_v0.txt > /dev/null 2> gpurun_out/r2_timeline_v0.err
head -3 gpurun_out/r2_timeline_v0.txt
python scripts/half_batch_probe.py > gpurun_out/r2_half_batch.txt 2>&1
cat gpurun_out/r2_half_batch.txt
( time python bench.py ) > gpurun_out/r2_bench_v0.json 2> gpurun_out/r2_bench_v0.err
tail -c 3000 gpurun_out/r2_bench_v0.json
tail -3 gpurun_out/r2_bench_v0.err
