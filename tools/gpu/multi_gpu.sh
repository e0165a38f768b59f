#!/bin/bash
# N-GPU validation of HEAD: bucket all-reduce test + bench.py under torchrun, N = 2 (default), 4 or 8.
# usage: gpurun --gpus N --timeout 1200 -- 'bash tools/gpu/multi_gpu.sh <tag> N'   (charged N x the box time)
mkdir -p gpurun_out
TAG=${1:-r1p}
N=${2:-2}
( timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -6 ) > gpurun_out/${TAG}_pytest_multi.txt 2>&1; tail -2 gpurun_out/${TAG}_pytest_multi.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${N} --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus ${N} --steps 4 --warmup 3 > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
python - <<'PY' ${TAG} ${N}
import json,sys
tag, n = sys.argv[1], sys.argv[2]
try:
    d=json.loads(open(f'gpurun_out/{tag}_bench_n{n}.json').read().strip().splitlines()[-1])
    r=d['roofline']; print(f'bench N={n} value',round(d['value'],2),'ms',round(d['ms_per_step'],1),'e2e',round(d['e2e']['value'],2),'gemmTF',round(r['achieved'],1),'share',round(r['gemm_share_of_step'],3),d['clocks'])
except Exception as e: print('bench failed',e, open(f'gpurun_out/{tag}_bench_n{n}.err').read()[-1500:])
PY
