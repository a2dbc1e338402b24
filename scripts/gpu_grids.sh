#!/bin/bash
# Grid-builder pass: parity tests under compute-sanitizer (small case) + plain, then a timing probe.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== sanitizer" ; timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_grids.py -x -q -k "single_point or unaligned" 2>&1 | tail -15 | tee gpurun_out/grids_sanitizer.log
echo "== pytest grids" ; timeout 900 python -m pytest tests/test_gpu_grids.py -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_grids.log
echo "== timing" ; timeout 600 python scripts/grids_once.py 2>&1 | tail -8 | tee gpurun_out/grids_once.log
