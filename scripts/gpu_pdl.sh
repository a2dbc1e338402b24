#!/bin/bash
# PDL on/off comparison of the graph-captured train step (+ parity tests with PDL on).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for v in 0 1 0 1; do
  echo "== NVF_PDL=$v" ; NVF_PDL=$v timeout 600 python bench.py --steps 40 --warmup 5 --skip-cpu-baseline --skip-epoch --skip-prep --decode-steps 2 > gpurun_out/bench_pdl$v.json 2> gpurun_out/bench_pdl$v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_pdl$v.json')); print('train', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'decode ms', round(d['decode']['ms_kernel'],3))"; tail -2 gpurun_out/bench_pdl$v.err
done
echo "== parity with PDL" ; NVF_PDL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
