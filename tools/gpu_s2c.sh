#!/bin/bash
# Session-2 call C: tcgen05 attention backward validation + timing (LLARK_ATTN_TC bit 0 = fwd, bit 1 = bwd).
mkdir -p gpurun_out
TAG=${1:-r1i}
( LLARK_ATTN_TC=3 timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k attention 2>&1 | tail -25 ) > gpurun_out/${TAG}_attn_tc_pytest.txt 2>&1
tail -12 gpurun_out/${TAG}_attn_tc_pytest.txt
LLARK_ATTN_TC=3 timeout 120 python tools/attn_time.py > gpurun_out/${TAG}_attn_time_tc.json 2>&1; tail -2 gpurun_out/${TAG}_attn_time_tc.json
LLARK_ATTN_TC=1 timeout 120 python tools/attn_time.py > gpurun_out/${TAG}_attn_time_tc1.json 2>&1; tail -1 gpurun_out/${TAG}_attn_time_tc1.json
if grep -q "passed" gpurun_out/${TAG}_attn_tc_pytest.txt && ! grep -q "[0-9] failed" gpurun_out/${TAG}_attn_tc_pytest.txt; then
  ( LLARK_ATTN_TC=3 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) > gpurun_out/${TAG}_pytest_tc_all.txt 2>&1; tail -1 gpurun_out/${TAG}_pytest_tc_all.txt
  for mode in 3 0; do
  LLARK_ATTN_TC=$mode timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_tc$mode.json 2> gpurun_out/${TAG}_bench_tc$mode.err
  python - <<'PY' ${TAG} $mode
import json,sys
try:
    d=json.loads(open(f'gpurun_out/{sys.argv[1]}_bench_tc{sys.argv[2]}.json').read().strip().splitlines()[-1])
    r=d['roofline']; print('bench tc mode',sys.argv[2],'B',d['config']['per_gpu_batch'],'value',round(d['value'],2),'ms',round(d['ms_per_step'],1),'gemmTF',round(r['achieved'],1),'share',round(r['gemm_share_of_step'],3),d['clocks'])
except Exception as e: print('bench_tc failed',e)
PY
  done
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"attn_bwd_d|attn_fwd_tc" -s 8 -c 3 -o gpurun_out/${TAG}_prof_attn_tc \
    env LLARK_ATTN_TC=3 python tools/attn_time.py > gpurun_out/${TAG}_ncu_attn.log 2>&1
fi
ls -la gpurun_out | tail -6
