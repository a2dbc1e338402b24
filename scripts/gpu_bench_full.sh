#!/bin/bash
# Bench (both arms) + ncu captures for profiles/: launch lists and full captures of the top kernels incl. the grid builder.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1; nproc >> gpurun_out/smi.txt
echo "== bench" ; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err ; tail -c 2500 gpurun_out/bench.json ; tail -5 gpurun_out/bench.err
echo "== bench reference arm" ; timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ; tail -c 1200 gpurun_out/bench_ref.json ; tail -3 gpurun_out/bench_ref.err
echo "== ncu grids launches + full" ; timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_edt_blocks|k_cells_insert|k_points_scatter' -s 6 -c 3 -o gpurun_out/prof_grids -f python scripts/grids_once.py --reps 1 > gpurun_out/ncu_grids.log 2>&1 ; tail -3 gpurun_out/ncu_grids.log
ls -la gpurun_out | head -40
