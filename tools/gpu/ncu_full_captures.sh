#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r1e}
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"adamw_kernel|attn_fwd_kernel|attn_bwd_dkdv|attn_bwd_dq|sumsq_partial" -s 6 -c 5 -o gpurun_out/${TAG}_prof_misc \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_misc.log 2>&1
timeout 1200 python tools/bench_jukebox.py --precision tf32 --steps 2 > gpurun_out/${TAG}_jukebox_tf32.json 2> gpurun_out/${TAG}_jukebox_tf32.err
tail -c 1500 gpurun_out/${TAG}_jukebox_tf32.json; tail -3 gpurun_out/${TAG}_jukebox_tf32.err
timeout 1200 python tools/bench_jukebox.py --precision bf16 --steps 2 > gpurun_out/${TAG}_jukebox_bf16.json 2> gpurun_out/${TAG}_jukebox_bf16.err
tail -c 1500 gpurun_out/${TAG}_jukebox_bf16.json; tail -3 gpurun_out/${TAG}_jukebox_bf16.err
ls -la gpurun_out | tail -8
