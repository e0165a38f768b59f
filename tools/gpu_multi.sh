#!/bin/bash
# 2-GPU validation of the data-parallel path (NCCL bucket all-reduce inside the library) + Jukebox bench.
mkdir -p gpurun_out
N=${1:-2}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 4 --warmup 2 > gpurun_out/multi_bench_$N.json 2> gpurun_out/multi_bench_$N.err
tail -c 1500 gpurun_out/multi_bench_$N.json; tail -5 gpurun_out/multi_bench_$N.err
