#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:'ba_|chain_|refit_|adaptive_|regen_best|rescale_points' -c 40 -o gpurun_out/prof_8f python tools/stages_8f.py > gpurun_out/ncu_8f.log 2>&1
tail -3 gpurun_out/ncu_8f.log | cut -c1-300
ls -la gpurun_out/prof_8f.ncu-rep
