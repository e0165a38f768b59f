#!/bin/bash
# A/B of the GEMM core: library built from commit 5e60258 vs the current one, same box, standalone shapes.
mkdir -p gpurun_out
TAG=${1:-r1m}
for i in 1 2; do
LLARK_LIB_PATH=$PWD/build_ab/libllark_b200_5e60258.so timeout 120 python tools/gemm_time.py 2>&1 | tail -1 | sed 's/^/old /' | tee -a gpurun_out/${TAG}_gemm_ab.txt
timeout 120 python tools/gemm_time.py 2>&1 | tail -1 | sed 's/^/new /' | tee -a gpurun_out/${TAG}_gemm_ab.txt
done
LLARK_LIB_PATH=$PWD/build_ab/libllark_b200_5e60258.so LLARK_FUSE_EPILOGUES=0 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_old.json 2> gpurun_out/${TAG}_bench_old.err
LLARK_FUSE_EPILOGUES=0 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_new.json 2> gpurun_out/${TAG}_bench_new.err
python - <<'PY' ${TAG}
import json,sys
for suf in ("old","new"):
    try:
        d=json.loads(open(f'gpurun_out/{sys.argv[1]}_bench_{suf}.json').read().strip().splitlines()[-1])
        r=d['roofline']; print('bench',suf,'value',round(d['value'],2),'ms',round(d['ms_per_step'],1),'gemmTF',round(r['achieved'],1),'share',round(r['gemm_share_of_step'],3),d['clocks'])
        for k,v in list(r['gemm_by_shape'].items())[:4]: print('    ',k,v)
    except Exception as e: print('bench',suf,'failed',e, open(f'gpurun_out/{sys.argv[1]}_bench_{suf}.err').read()[-500:])
PY
