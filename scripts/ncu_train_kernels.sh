#!/bin/bash
# Full ncu capture of selected training kernels (one launch each) from one train step.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PAT="${1:-k_conv_s1|k_wgrad4_s1}"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:$PAT" -o gpurun_out/prof_train -f python scripts/train_once.py --no-graph > gpurun_out/ncu_train.log 2>&1
tail -3 gpurun_out/ncu_train.log
ls -la gpurun_out/prof_train.ncu-rep
