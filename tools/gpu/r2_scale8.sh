#!/bin/bash
# Round-2 scaling pass on ONE 8-GPU box (gpurun --gpus 8): same-box N = 1 and N = 8 numbers of the default schedule, one
# ablation of the round-2 scaling changes, the N > 1 correctness test, and configs[4] on 2 GPUs.
# usage: gpurun --gpus 8 --timeout 1500 -- 'bash tools/gpu/r2_scale8.sh r2s'
mkdir -p gpurun_out
TAG=${1:-r2s}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
Q="--steps 4 --warmup 3 --no-cpu-baseline --no-torch-baseline"
nvidia-smi -L | head -8
# N = 8, default schedule, with the configs[3] extra; NCCL's own choice of algorithm logged once
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 600 $TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 $Q \
  > gpurun_out/${TAG}_n8_default.json 2> gpurun_out/${TAG}_n8_default.err
grep -E "NVLS|Using network|Channel [0-9]+/[0-9]+|algo|Connected all" gpurun_out/${TAG}_n8_default.err | sort | uniq -c | sort -rn | head -12 > gpurun_out/${TAG}_nccl_info.txt
tail -c 20000 gpurun_out/${TAG}_n8_default.err > gpurun_out/${TAG}_n8_default.err.tail; rm -f gpurun_out/${TAG}_n8_default.err
# same-box N = 1 (efficiency denominator)
timeout 400 python bench.py --gpus 1 $Q --no-extras > gpurun_out/${TAG}_n1.json 2>/dev/null
# ablation at N = 8: tail split off (equal 4-layer slices, as in round 1)
LLARK_COMM_TAIL_SPLIT=0 timeout 400 $TR --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 $Q --no-extras \
  > gpurun_out/${TAG}_n8_no_tail_split.json 2>/dev/null
# N > 1 correctness (2 GPUs) and configs[4] on 2 GPUs
( timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -3 ) > gpurun_out/${TAG}_pytest_multi.txt
timeout 300 $TR --nproc-per-node 2 --master-port 29514 tools/bench_mpt.py --steps 5 --warmup 3 \
  > gpurun_out/${TAG}_mpt_n2.json 2>/dev/null
for f in n8_default n1 n8_no_tail_split; do python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_$f.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$f", "value", round(d["value"], 2), "ms", round(d["ms_per_step"], 1), "gemm", round(r["achieved"], 1),
          "checks", d.get("multi_gpu_checks"), "c3", (d.get("configs3_e2e") or {}).get("value"))
except Exception as e:
    print("$f failed", e)
PY
done
cat gpurun_out/${TAG}_pytest_multi.txt; cut -c1-300 gpurun_out/${TAG}_mpt_n2.json; cat gpurun_out/${TAG}_nccl_info.txt
