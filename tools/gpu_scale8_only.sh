#!/bin/bash
# 8-GPU point of the scaling table only (bench pairs-sharded, configs 3 and 4); run with gpurun --gpus 8
mkdir -p gpurun_out
N=8
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29608"
echo "== bench N=$N"; timeout 600 $RUN bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/scale_bench_n$N.json 2> gpurun_out/scale_bench_n$N.err; tail -1 gpurun_out/scale_bench_n$N.json | cut -c1-220
echo "== configs N=$N"; timeout 900 $RUN tools/configs.py c3 c4 > gpurun_out/scale_configs_n$N.jsonl 2> gpurun_out/scale_configs_n$N.err; cut -c1-260 gpurun_out/scale_configs_n$N.jsonl
