#!/bin/bash
# First on-device pass: parity tests, smoke, bench, launch list, one full ncu capture of the fused decode kernel.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench" ; timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err ; tail -c 3000 gpurun_out/bench.json ; tail -5 gpurun_out/bench.err
echo "== ncu launches" ; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --decode-steps 1 --skip-cpu-baseline --train-blocks 16 > gpurun_out/ncu_bench.log 2>&1 ; tail -2 gpurun_out/ncu_bench.log
echo "== ncu full" ; timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_decode_fused -c 1 -o gpurun_out/prof_fusedA -f python scripts/decode_once.py > gpurun_out/ncu_full.log 2>&1 ; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
