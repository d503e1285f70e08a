#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; tail -40 gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== diag" ; timeout 600 python tests/diag/gpu_diag.py 2>&1 | tee gpurun_out/diag.log
echo "== probe" ; timeout 900 python tools/gpu_probe.py 2>&1 | tee gpurun_out/probe.jsonl | cut -c1-900
ls -la gpurun_out
