#!/bin/bash
# N = 2 validation of the data-parallel path (gpurun --gpus 2): bench.py under torchrun (multi_gpu_checks in the line),
# tests/test_multi_gpu.py, configs[4] (MPT-1B) on 2 GPUs.
mkdir -p gpurun_out
TAG=${1:-r2n2}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 500 $TR --nproc-per-node 2 --master-port 29521 bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu-baseline --no-torch-baseline --no-extras \
  > gpurun_out/${TAG}_bench.out 2> gpurun_out/${TAG}_bench.err
grep '^{"metric"' gpurun_out/${TAG}_bench.out | tail -1 > gpurun_out/${TAG}_bench.json
( timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -3 ) > gpurun_out/${TAG}_pytest_multi.txt
timeout 300 $TR --nproc-per-node 2 --master-port 29524 tools/bench_mpt.py --steps 5 --warmup 3 \
  > gpurun_out/${TAG}_mpt.out 2> gpurun_out/${TAG}_mpt.err
grep '^{"workload"' gpurun_out/${TAG}_mpt.out | tail -1 > gpurun_out/${TAG}_mpt.json
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench.json").read())
    print("N=2 value", round(d["value"], 2), "ms", round(d["ms_per_step"], 1), "gemm", round(d["roofline"]["achieved"], 1), d.get("multi_gpu_checks"))
except Exception as e:
    print("bench failed", e)
PY
tail -3 gpurun_out/${TAG}_bench.err | cut -c1-300
cat gpurun_out/${TAG}_pytest_multi.txt; cut -c1-300 gpurun_out/${TAG}_mpt.json; tail -2 gpurun_out/${TAG}_mpt.err | cut -c1-300
