#!/bin/bash
# Iteration pass: parity tests + launch list of one train step + bench without the CPU baseline.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== ncu train launches" ; timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/launches_train.csv python scripts/train_once.py "$@" > gpurun_out/train_once.log 2>&1 ; tail -2 gpurun_out/train_once.log
echo "== bench" ; timeout 900 python bench.py --steps 20 --warmup 3 --skip-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err ; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print('train', d['value'], d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['value']); print('decode', d['decode']['value'], d['decode']['ms_kernel'], d['decode']['roofline']['frac'])" ; tail -5 gpurun_out/bench.err
