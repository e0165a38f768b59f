#!/bin/bash
# Session-2 call F: templated GEMM epilogues (plain GEMM code unchanged), double-buffered attention backward, batched conv / attention tile loads.
mkdir -p gpurun_out
TAG=${1:-r1l}
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > gpurun_out/${TAG}_pytest.txt 2>&1; tail -3 gpurun_out/${TAG}_pytest.txt
timeout 120 python tools/attn_time.py > gpurun_out/${TAG}_attn_time.json 2>&1; tail -1 gpurun_out/${TAG}_attn_time.json
timeout 120 python tools/gemm_time.py > gpurun_out/${TAG}_gemm_time.txt 2>&1; tail -2 gpurun_out/${TAG}_gemm_time.txt
for fuse in 1 0; do
LLARK_FUSE_EPILOGUES=$fuse timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_fuse$fuse.json 2> gpurun_out/${TAG}_bench_fuse$fuse.err
python - <<'PY' ${TAG} $fuse
import json,sys
try:
    d=json.loads(open(f'gpurun_out/{sys.argv[1]}_bench_fuse{sys.argv[2]}.json').read().strip().splitlines()[-1])
    r=d['roofline']; print('bench fuse',sys.argv[2],'B',d['config']['per_gpu_batch'],'value',round(d['value'],2),'ms',round(d['ms_per_step'],1),'e2e',round(d['e2e']['value'],2),'gemmTF',round(r['achieved'],1),'share',round(r['gemm_share_of_step'],3),d['clocks'])
except Exception as e: print('bench failed',e)
PY
done
timeout 600 python tools/bench_jukebox.py --precision bf16 --steps 2 > gpurun_out/${TAG}_jukebox_bf16.json 2> gpurun_out/${TAG}_jukebox_bf16.err
tail -c 900 gpurun_out/${TAG}_jukebox_bf16.json; tail -2 gpurun_out/${TAG}_jukebox_bf16.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_jukebox_launches.csv \
  python tools/bench_jukebox.py --precision bf16 --steps 1 --warmup 0 --depth 3 --batch 8 > gpurun_out/${TAG}_ncu_jb.log 2>&1
python tools/summarize_ncu.py launches gpurun_out/${TAG}_jukebox_launches.csv gpurun_out/${TAG}_jukebox_launches_summary.md 2>&1 | head -22
