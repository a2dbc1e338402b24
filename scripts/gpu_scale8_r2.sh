#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29591"
( time NVF_PEER_ALLREDUCE=1 timeout 900 $T bench.py --gpus 8 --steps 20 --warmup 3 --skip-cpu-baseline ) > gpurun_out/r2_bench_n8_peer1.json 2> gpurun_out/r2_bench_n8_peer1.err
tail -c 2600 gpurun_out/r2_bench_n8_peer1.json; echo
NVF_PEER_ALLREDUCE=0 timeout 600 $T bench.py --gpus 8 --steps 20 --warmup 3 --skip-prep --skip-wide --skip-epoch --skip-cpu-baseline 2> gpurun_out/r2_bench_n8_peer0.err | head -c 330; echo
grep -i "peer-memory" gpurun_out/r2_bench_n8_peer*.err | head
