#!/bin/bash
# Final pass of the round: full GPU test tier, smoke, both bench arms, and the ncu launch list of the bench command itself.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== bench" ; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err ; tail -c 600 gpurun_out/bench.json ; tail -3 gpurun_out/bench.err
echo "== bench reference arm" ; timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ; tail -c 300 gpurun_out/bench_ref.json ; tail -3 gpurun_out/bench_ref.err
echo "== ncu launch list of the bench command" ; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1500 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --skip-cpu-baseline --skip-epoch --decode-steps 1 > gpurun_out/bench_under_ncu.log 2>&1 ; tail -c 200 gpurun_out/bench_under_ncu.log
ls -la gpurun_out | head -30
