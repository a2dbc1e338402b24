#!/bin/bash
# 8-GPU pass: default bench (vox10, config A) and the vox11 / wide-channel configuration (BASELINE.json configs[3-4]).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N="${1:-8}"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N"
echo "== bench N=$N vox10 A" ; timeout 300 $TR --steps 20 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_n${N}_A.json 2> gpurun_out/bench_n${N}_A.err ; tail -c 400 gpurun_out/bench_n${N}_A.json ; tail -2 gpurun_out/bench_n${N}_A.err
echo "== bench N=$N vox11 B" ; timeout 300 $TR --steps 10 --warmup 3 --skip-cpu-baseline --skip-epoch --skip-prep --chanstr 16,32,16,16 --resolution 2048 --decode-steps 3 > gpurun_out/bench_n${N}_B.json 2> gpurun_out/bench_n${N}_B.err ; tail -c 400 gpurun_out/bench_n${N}_B.json ; tail -2 gpurun_out/bench_n${N}_B.err
