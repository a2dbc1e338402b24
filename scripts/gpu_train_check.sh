#!/bin/bash
# quick check after a train-path change: parity tests, bench line, concurrent timeline of one step
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py tests/test_gpu_step.py tests/test_gpu_fit.py -x -q -k "train or step or fit or adam or weight or feeder" ) 2>&1 | tail -2
python bench.py --skip-prep --skip-wide --skip-epoch --skip-cpu-baseline --steps 30 2>/dev/null | head -c 230; echo
python scripts/timeline.py --out gpurun_out/r2_timeline_latest.txt > /dev/null 2>&1
head -1 gpurun_out/r2_timeline_latest.txt
