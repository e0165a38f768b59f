#!/bin/bash
# Session-2 call D: register-tiled Jukebox conv kernels (tests + bench), prior launch breakdown, ncu of the tcgen05 backward kernels.
mkdir -p gpurun_out
TAG=${1:-r1j}
( timeout 600 python -m pytest tests/test_jukebox_gpu.py -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/${TAG}_jukebox_pytest.txt 2>&1; tail -2 gpurun_out/${TAG}_jukebox_pytest.txt
timeout 600 python tools/bench_jukebox.py --precision bf16 --steps 2 > gpurun_out/${TAG}_jukebox_bf16.json 2> gpurun_out/${TAG}_jukebox_bf16.err
tail -c 900 gpurun_out/${TAG}_jukebox_bf16.json; tail -2 gpurun_out/${TAG}_jukebox_bf16.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_jukebox_launches.csv \
  python tools/bench_jukebox.py --precision bf16 --steps 1 --warmup 0 --depth 3 --batch 8 > gpurun_out/${TAG}_ncu_jb.log 2>&1
python tools/summarize_ncu.py launches gpurun_out/${TAG}_jukebox_launches.csv gpurun_out/${TAG}_jukebox_launches_summary.md 2>&1 | head -30
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"attn_bwd_d.*_tc" -s 4 -c 2 -o gpurun_out/${TAG}_prof_attn_bwd_tc \
  python tools/attn_time.py > gpurun_out/${TAG}_ncu_attn.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"resblock" -s 2 -c 2 -o gpurun_out/${TAG}_prof_resblock \
  python tools/bench_jukebox.py --precision bf16 --steps 1 --warmup 0 --depth 1 --batch 2 > gpurun_out/${TAG}_ncu_res.log 2>&1
ls -la gpurun_out | tail -6
