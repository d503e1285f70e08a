#!/bin/bash
# First GPU call: tests, probes, bench (both arms), ncu launch list + full capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | tail -80 | tee gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== probe" ; timeout 600 python tools/gpu_probe.py 2>&1 | tee gpurun_out/probe.jsonl | cut -c1-600
echo "== bench" ; timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tee gpurun_out/bench.json | cut -c1-3000
echo "== bench reference" ; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tee gpurun_out/bench_ref.json | cut -c1-2000
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log | cut -c1-300
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_kernel -s 3 -c 2 -o gpurun_out/prof_score python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out
