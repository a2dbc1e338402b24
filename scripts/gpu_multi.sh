#!/bin/bash
# Multi-GPU pass: bench at N ranks (default 2) + the reference arm launch path.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N="${1:-2}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | tee gpurun_out/smi_multi.txt
echo "== bench N=$N" ; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err ; tail -c 3000 gpurun_out/bench_n$N.json ; tail -5 gpurun_out/bench_n$N.err
echo "== bench N=1" ; timeout 900 python bench.py --steps 20 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err ; tail -c 600 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
