#!/bin/bash
# Session-2 verification of HEAD: GPU tests + smoke + bench with the dynamic and the static GEMM tile scheduler + ncu launch list.
mkdir -p gpurun_out
TAG=${1:-r1g}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -30 ) > gpurun_out/${TAG}_pytest.txt 2>&1
tail -4 gpurun_out/${TAG}_pytest.txt
if ! grep -q " passed" gpurun_out/${TAG}_pytest.txt || grep -q "failed" gpurun_out/${TAG}_pytest.txt; then
  echo "--- rerun with static tile scheduler (LLARK_GEMM_DEBUG=32)"
  ( LLARK_GEMM_DEBUG=32 timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/${TAG}_pytest_static.txt 2>&1
  tail -4 gpurun_out/${TAG}_pytest_static.txt
fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1; tail -2 gpurun_out/${TAG}_smoke.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 2500 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
LLARK_GEMM_DEBUG=32 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_static.json 2> gpurun_out/${TAG}_bench_static.err
python - <<'PY' ${TAG}
import json,sys
for suf in ("bench","bench_static"):
    try:
        d=json.loads(open(f'gpurun_out/{sys.argv[1]}_{suf}.json').read().strip().splitlines()[-1])
        r=d['roofline']; print(suf,'B',d['config']['per_gpu_batch'],'value',round(d['value'],2),'ms',round(d['ms_per_step'],1),'e2e',round(d['e2e']['value'],2),'gemmTF',round(r['achieved'],1),'share',round(r['gemm_share_of_step'],3),d['clocks'])
    except Exception as e: print(suf,'failed',e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
python tools/summarize_ncu.py launches gpurun_out/${TAG}_launches.csv gpurun_out/${TAG}_launches_summary.md 2>&1 | head -32
