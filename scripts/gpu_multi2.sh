#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
( timeout 900 python -m pytest tests/test_gpu_multi.py -x -q ) > gpurun_out/r2_multi_pytest.log 2>&1
tail -15 gpurun_out/r2_multi_pytest.log
for peer in 1 0; do
  NVF_PEER_ALLREDUCE=$peer timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --steps 20 --warmup 3 --skip-prep --skip-wide --skip-epoch --skip-cpu-baseline 2> gpurun_out/r2_bench_n2_peer$peer.err | head -c 330; echo
  grep -i "peer-memory\|error" gpurun_out/r2_bench_n2_peer$peer.err | head -3
done
