#!/bin/bash
# Iteration pass: parity tests + launch list of one train step + bench.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== ncu train launches" ; timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/launches_train.csv python scripts/train_once.py "$@" > gpurun_out/train_once.log 2>&1 ; tail -2 gpurun_out/train_once.log
echo "== bench" ; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err ; tail -c 2500 gpurun_out/bench.json ; tail -5 gpurun_out/bench.err
