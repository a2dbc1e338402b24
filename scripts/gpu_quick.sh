#!/bin/bash
# Quick on-device pass: parity tests + bench (optionally skipping the CPU baseline).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== bench" ; timeout 900 python bench.py --steps 20 --warmup 3 "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err ; tail -c 4000 gpurun_out/bench.json ; tail -5 gpurun_out/bench.err
