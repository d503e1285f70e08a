#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
for cfg in 128,2,0 128,2,1 128,2,2 256,1,0 256,1,1 256,1,2 64,4,0 64,4,1 64,4,2; do
  echo "== hypgen $cfg"; SFMB200_HYPGEN=$cfg timeout 300 python tools/gpu_probe.py hypgen > gpurun_out/probe_hypgen_$cfg.jsonl 2>&1
done
echo "== c5"; timeout 600 python tools/configs.py c5 2>&1 | tail -1 | cut -c1-700
ls gpurun_out
