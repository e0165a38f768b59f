#!/bin/bash
# 2-GPU validation of HEAD: bucket all-reduce test + bench.py under torchrun.
mkdir -p gpurun_out
TAG=${1:-r1p}
( timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -6 ) > gpurun_out/${TAG}_pytest_multi.txt 2>&1; tail -2 gpurun_out/${TAG}_pytest_multi.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err
python - <<'PY' ${TAG}
import json,sys
try:
    d=json.loads(open(f'gpurun_out/{sys.argv[1]}_bench_n2.json').read().strip().splitlines()[-1])
    r=d['roofline']; print('bench N=2 value',round(d['value'],2),'ms',round(d['ms_per_step'],1),'e2e',round(d['e2e']['value'],2),'gemmTF',round(r['achieved'],1),'share',round(r['gemm_share_of_step'],3),d['clocks'])
except Exception as e: print('bench failed',e, open(f'gpurun_out/{sys.argv[1]}_bench_n2.err').read()[-1500:])
PY
