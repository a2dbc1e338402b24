#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pytest gpu (PDL on)" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for v in 1 0 1 0; do
  echo "== bench NVF_PDL=$v"
  NVF_PDL=$v timeout 600 python bench.py --steps 30 --warmup 3 --skip-cpu-baseline --skip-epoch > gpurun_out/bench_pdl$v.json 2> gpurun_out/bench_pdl$v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_pdl$v.json')); print('train', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'decode', d['decode']['ms_per_step'], d['decode']['ms_kernel'])"; tail -2 gpurun_out/bench_pdl$v.err
done
