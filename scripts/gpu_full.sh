#!/bin/bash
# Full on-device pass: parity tests, smoke, bench, launch lists (train step + decode), full ncu capture of
# the fused decode kernel and of the top train kernels.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench" ; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err ; tail -c 3500 gpurun_out/bench.json ; tail -5 gpurun_out/bench.err
echo "== ncu train launches" ; timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/launches_train.csv python scripts/train_once.py > gpurun_out/train_once.log 2>&1 ; tail -2 gpurun_out/train_once.log
echo "== ncu decode launches" ; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_decode|k_emit|k_scan|k_' --csv --log-file gpurun_out/launches_decode.csv python scripts/decode_once.py > gpurun_out/decode_once.log 2>&1 ; tail -2 gpurun_out/decode_once.log
echo "== ncu full decode" ; timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_decode_fused -c 1 -o gpurun_out/prof_fusedA -f python scripts/decode_once.py > gpurun_out/ncu_full.log 2>&1 ; tail -3 gpurun_out/ncu_full.log
echo "== ncu full train" ; timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:k_conv_s1|k_wgrad4_s1|k_stem_bwd|k_cls_wgrad|k_convT5" -o gpurun_out/prof_train -f python scripts/train_once.py --no-graph > gpurun_out/ncu_train.log 2>&1 ; tail -3 gpurun_out/ncu_train.log
ls -la gpurun_out
