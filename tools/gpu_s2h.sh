#!/bin/bash
# Session-2 call H: lean TMA producer loop (A/B vs the 5e60258 build), full GPU tests, bench with fused epilogues.
mkdir -p gpurun_out
TAG=${1:-r1o}
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > gpurun_out/${TAG}_pytest.txt 2>&1; tail -2 gpurun_out/${TAG}_pytest.txt
for i in 1 2; do
LLARK_LIB_PATH=$PWD/build_ab/libllark_b200_5e60258.so timeout 120 python tools/gemm_time.py 2>&1 | tail -1 | sed 's/^/old /' | tee -a gpurun_out/${TAG}_gemm_ab.txt
timeout 120 python tools/gemm_time.py 2>&1 | tail -1 | sed 's/^/new /' | tee -a gpurun_out/${TAG}_gemm_ab.txt
done
timeout 120 python tools/attn_time.py > gpurun_out/${TAG}_attn_time.json 2>&1; tail -1 gpurun_out/${TAG}_attn_time.json
for fuse in 1 0; do
LLARK_FUSE_EPILOGUES=$fuse timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_fuse$fuse.json 2> gpurun_out/${TAG}_bench_fuse$fuse.err
python - <<'PY' ${TAG} $fuse
import json,sys
try:
    d=json.loads(open(f'gpurun_out/{sys.argv[1]}_bench_fuse{sys.argv[2]}.json').read().strip().splitlines()[-1])
    r=d['roofline']; print('bench fuse',sys.argv[2],'value',round(d['value'],2),'ms',round(d['ms_per_step'],1),'e2e',round(d['e2e']['value'],2),'gemmTF',round(r['achieved'],1),'share',round(r['gemm_share_of_step'],3),d['clocks'])
    for k,v in list(r['gemm_by_shape'].items())[:3]: print('    ',k,v)
except Exception as e: print('bench failed',e)
PY
done
