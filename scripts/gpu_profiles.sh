#!/bin/bash
# Profile pass for profiles/: launch lists (train step, decode) + full ncu captures of the top kernels.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== ncu train launches" ; timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/launches_train.csv python scripts/train_once.py > gpurun_out/train_once.log 2>&1 ; tail -1 gpurun_out/train_once.log
echo "== ncu decode launches" ; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_' --csv --log-file gpurun_out/launches_decode.csv python scripts/decode_once.py > gpurun_out/decode_once.log 2>&1 ; tail -1 gpurun_out/decode_once.log
echo "== ncu full decode (the 1247-block launch)" ; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_decode_fused -s 1 -c 1 -o gpurun_out/prof_fusedA -f python scripts/decode_once.py > gpurun_out/ncu_full.log 2>&1 ; tail -2 gpurun_out/ncu_full.log
echo "== ncu full train" ; timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:k_conv_s1<4|k_wgrad4_s1|k_convT5" -o gpurun_out/prof_train -f python scripts/train_once.py --no-graph > gpurun_out/ncu_train.log 2>&1 ; tail -2 gpurun_out/ncu_train.log
ls -la gpurun_out | head -30
