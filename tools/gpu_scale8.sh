#!/bin/bash
# 1/2/4/8-GPU scaling on one box: bench (pairs sharded) and configs 3/4 (run with gpurun --gpus 8)
mkdir -p gpurun_out
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then RUN="python"; else RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N))"; fi
  echo "== bench N=$N"; timeout 600 $RUN bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/scale_bench_n$N.json 2> gpurun_out/scale_bench_n$N.err; cut -c1-220 gpurun_out/scale_bench_n$N.json
  echo "== configs N=$N"; timeout 900 $RUN tools/configs.py c3 c4 > gpurun_out/scale_configs_n$N.jsonl 2> gpurun_out/scale_configs_n$N.err; cut -c1-260 gpurun_out/scale_configs_n$N.jsonl
done
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
