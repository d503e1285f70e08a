#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== diag" ; timeout 600 python tests/diag/gpu_diag.py 2>&1 | head -8 | tee gpurun_out/diag.log
echo "== probe" ; timeout 900 python tools/gpu_probe.py > gpurun_out/probe.jsonl 2>&1; tail -2 gpurun_out/probe.jsonl | cut -c1-300
for m in 2 3 4; do SFMB200_HYPGEN_MINB=$m timeout 300 python tools/gpu_probe.py hypgen > gpurun_out/probe_hypgen_$m.jsonl 2>&1; done
ls -la gpurun_out
