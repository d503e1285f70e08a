#!/bin/bash
# 2-GPU validation: bench (pairs sharded), configs 3/4/5; run with gpurun --gpus 2
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
echo "== facade + gpu tests" ; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
echo "== bench N=1" ; timeout 600 python bench.py > gpurun_out/bench_n1.json 2>gpurun_out/bench_n1.err; cut -c1-400 gpurun_out/bench_n1.json; tail -2 gpurun_out/bench_n1.err
echo "== bench N=$N" ; timeout 600 $TR bench.py --gpus $N > gpurun_out/bench_n$N.json 2>gpurun_out/bench_n$N.err; cut -c1-400 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
echo "== bench reference N=$N" ; timeout 900 $TR bench.py --gpus $N --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_n$N.json 2>gpurun_out/bench_ref_n$N.err; cut -c1-300 gpurun_out/bench_ref_n$N.json; tail -2 gpurun_out/bench_ref_n$N.err
echo "== configs N=1" ; timeout 900 python tools/configs.py c3 c4 c5 > gpurun_out/configs_n1.jsonl 2>gpurun_out/configs_n1.err; cut -c1-500 gpurun_out/configs_n1.jsonl; tail -2 gpurun_out/configs_n1.err
echo "== configs N=$N" ; timeout 900 $TR tools/configs.py c3 c4 c5 > gpurun_out/configs_n$N.jsonl 2>gpurun_out/configs_n$N.err; cut -c1-500 gpurun_out/configs_n$N.jsonl; tail -3 gpurun_out/configs_n$N.err
ls gpurun_out
