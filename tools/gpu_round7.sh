#!/bin/bash
# round-1 late pass: whole GPU suite + timings of the 8f additions
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r7_tests.txt
python tools/ba_timing.py > gpurun_out/ba_timing.jsonl 2> gpurun_out/ba_timing.err
python bench.py --steps 300 --warmup 5 > gpurun_out/r7_bench.json 2> gpurun_out/r7_bench.err
cat gpurun_out/r7_tests.txt; cat gpurun_out/ba_timing.jsonl; tail -3 gpurun_out/ba_timing.err; cut -c1-300 gpurun_out/r7_bench.json
