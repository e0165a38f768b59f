#!/bin/bash
# Session-2 call B: tcgen05 attention forward validation + timing, Jukebox front-end bench (configs[1]) + ncu captures.
mkdir -p gpurun_out
TAG=${1:-r1h}
( LLARK_ATTN_TC=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k attention 2>&1 | tail -15 ) > gpurun_out/${TAG}_attn_tc_pytest.txt 2>&1
tail -3 gpurun_out/${TAG}_attn_tc_pytest.txt
LLARK_ATTN_TC=0 timeout 120 python tools/attn_time.py > gpurun_out/${TAG}_attn_time_mma.json 2>&1; tail -1 gpurun_out/${TAG}_attn_time_mma.json
LLARK_ATTN_TC=1 timeout 120 python tools/attn_time.py > gpurun_out/${TAG}_attn_time_tc.json 2>&1; tail -2 gpurun_out/${TAG}_attn_time_tc.json
if grep -q "passed" gpurun_out/${TAG}_attn_tc_pytest.txt && ! grep -q "[0-9] failed" gpurun_out/${TAG}_attn_tc_pytest.txt; then
  ( LLARK_ATTN_TC=1 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) > gpurun_out/${TAG}_pytest_tc_all.txt 2>&1; tail -1 gpurun_out/${TAG}_pytest_tc_all.txt
  LLARK_ATTN_TC=1 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_tc.json 2> gpurun_out/${TAG}_bench_tc.err
  python - <<'PY' ${TAG}
import json,sys
try:
    d=json.loads(open(f'gpurun_out/{sys.argv[1]}_bench_tc.json').read().strip().splitlines()[-1])
    r=d['roofline']; print('bench_tc B',d['config']['per_gpu_batch'],'value',round(d['value'],2),'ms',round(d['ms_per_step'],1),'gemmTF',round(r['achieved'],1),'share',round(r['gemm_share_of_step'],3),d['clocks'])
except Exception as e: print('bench_tc failed',e)
PY
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc -s 4 -c 2 -o gpurun_out/${TAG}_prof_attn_tc \
    env LLARK_ATTN_TC=1 python tools/attn_time.py > gpurun_out/${TAG}_ncu_attn.log 2>&1
fi
timeout 600 python tools/bench_jukebox.py --precision bf16 --steps 2 > gpurun_out/${TAG}_jukebox_bf16.json 2> gpurun_out/${TAG}_jukebox_bf16.err
tail -c 1200 gpurun_out/${TAG}_jukebox_bf16.json; tail -2 gpurun_out/${TAG}_jukebox_bf16.err
timeout 600 python tools/bench_jukebox.py --precision tf32 --steps 2 > gpurun_out/${TAG}_jukebox_tf32.json 2> gpurun_out/${TAG}_jukebox_tf32.err
tail -c 1200 gpurun_out/${TAG}_jukebox_tf32.json; tail -2 gpurun_out/${TAG}_jukebox_tf32.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"resblock|conv_down|conv_k3|vq_encode|jb_pool|jb_sparse_attn|layernorm|quick_gelu" -c 24 -o gpurun_out/${TAG}_prof_jukebox \
  python tools/bench_jukebox.py --precision bf16 --steps 1 --warmup 0 --depth 2 --batch 2 > gpurun_out/${TAG}_ncu_jukebox.log 2>&1
ls -la gpurun_out | tail -12
