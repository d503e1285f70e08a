#!/bin/bash
# Official-style bench (both arms) + ncu evidence for profiles/.
mkdir -p gpurun_out
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
echo "== bench" ; timeout 600 python bench.py > gpurun_out/bench.json 2>gpurun_out/bench.err; cut -c1-1500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== bench reference" ; timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cut -c1-1200 gpurun_out/bench_ref.json
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
grep -c "sfmb200" gpurun_out/launches.csv
echo "== ncu full (score, hypgen, triangulate, fused)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'score_kernel|hypgen_kernel|triangulate_kernel|select_pose' -s 12 -c 8 -o gpurun_out/prof_path python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
ls -la gpurun_out
