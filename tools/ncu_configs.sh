#!/bin/bash
# ncu --set full of the scoring kernel at the config-3 and config-4 shapes and of hypgen / triangulation / the fused small
# path; reports land in gpurun_out/ (summarise with tools/ncu_summary.py into profiles/).
set -x
cd "$(dirname "$0")/.."
ncu --set full --clock-control none --import-source on -k regex:score_kernel -c 1 -o gpurun_out/r02_score_c3 python tools/configs.py c3 --reps 1 > gpurun_out/ncu_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:score_kernel -c 1 -o gpurun_out/r02_score_c4 python tools/configs.py c4 --scale 0.25 --reps 1 > gpurun_out/ncu_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hypgen_kernel -c 1 -o gpurun_out/r02_hypgen_c4 python tools/configs.py c4 --scale 0.25 --reps 1 > gpurun_out/ncu_c4h.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"score_kernel|hypgen_kernel|triangulate_kernel|select_pose|ingest_xy" -c 12 -o gpurun_out/r02_bench_c2 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:triangulate -c 1 -o gpurun_out/r02_tri3 python tools/tri_one.py > gpurun_out/tri_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:small_path -c 1 -o gpurun_out/r02_small2 python tools/small_phases.py > gpurun_out/small_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
