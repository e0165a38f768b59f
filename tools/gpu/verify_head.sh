#!/bin/bash
# Session-2 call J: 6-slot rings in the attention backward kernels; full tests, bench, ncu launch list of the step.
mkdir -p gpurun_out
TAG=${1:-r1q}
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > gpurun_out/${TAG}_pytest.txt 2>&1; tail -2 gpurun_out/${TAG}_pytest.txt
timeout 120 python tools/attn_time.py > gpurun_out/${TAG}_attn_time.json 2>&1; tail -1 gpurun_out/${TAG}_attn_time.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1; tail -1 gpurun_out/${TAG}_smoke.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<'PY' ${TAG}
import json,sys
try:
    d=json.loads(open(f'gpurun_out/{sys.argv[1]}_bench.json').read().strip().splitlines()[-1])
    r=d['roofline']; print('bench value',round(d['value'],2),'ms',round(d['ms_per_step'],1),'e2e',round(d['e2e']['value'],2),'gemmTF',round(r['achieved'],1),'share',round(r['gemm_share_of_step'],3),d['clocks'], d.get('cpu_baseline',{}).get('value'))
except Exception as e: print('bench failed',e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
python tools/summarize_ncu.py launches gpurun_out/${TAG}_launches.csv gpurun_out/${TAG}_launches_summary.md 2>&1 | head -34
