#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r1f}
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/${TAG}_pytest.txt; tail -4 gpurun_out/${TAG}_pytest.txt
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<'PY'
import json,sys
d=json.loads(open('gpurun_out/'+sys.argv[1]+'_bench.json' if len(sys.argv)>1 else 'gpurun_out/r1f_bench.json').read().strip().splitlines()[-1])
r=d['roofline']; print('B',d['config']['per_gpu_batch'],'value',round(d['value'],2),'ms',round(d['ms_per_step'],1),'e2e',round(d['e2e']['value'],2),'gemmTF',round(r['achieved'],1),'share',round(r['gemm_share_of_step'],3),d['clocks'])
PY
timeout 1500 python bench.py --steps 5 --warmup 3 --batch 8 --no-cpu-baseline > gpurun_out/${TAG}_bench_b8.json 2> gpurun_out/${TAG}_bench_b8.err
tail -c 400 gpurun_out/${TAG}_bench_b8.json
