#!/bin/bash
# end-of-round evidence: GPU tier, default bench lines, launch lists, ncu --set full of the decode head kernels
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r2_pytest_gpu.log 2>&1
tail -4 gpurun_out/r2_pytest_gpu.log | head -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
tail -c 600 gpurun_out/r2_bench_n1.json; echo
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
head -c 400 gpurun_out/r2_bench_reference.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --skip-prep --skip-wide --skip-epoch --skip-cpu-baseline > gpurun_out/r2_launches_bench.log 2>&1
for cfg in 8,16,8,8 16,32,16,16; do
  tag=$(echo $cfg | tr , _)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_decode_head -s 2 -c 1 -o gpurun_out/r2_head_$tag python scripts/decode_probe.py $cfg 1024 1 > gpurun_out/r2_head_ncu_$tag.log 2>&1
done
ls gpurun_out/*.ncu-rep
