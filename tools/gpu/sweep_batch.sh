#!/bin/bash
# Per-GPU micro-batch sweep of the headline step on one box (same clocks, same power cap): the optimizer is paid once per
# micro-batch in bench.py, so larger batches amortise its ~16 ms; 54 GB of state + 2.16 GB/sample of activations must fit.
# usage: gpurun --timeout 1200 -- 'bash tools/gpu/sweep_batch.sh r2b "16 24 32"'
mkdir -p gpurun_out
TAG=${1:-r2b}
for B in ${2:-16 24 32}; do
  timeout 400 python bench.py --steps 4 --warmup 3 --batch $B --no-cpu-baseline > gpurun_out/${TAG}_bench_b${B}.json 2> gpurun_out/${TAG}_bench_b${B}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_b${B}.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("B=${B}", "samples/s", round(d["value"], 2), "ms/step", round(d["ms_per_step"], 1), "gemm TF/s", round(r["achieved"], 1),
          "gemm share", round(r["gemm_share_of_step"], 3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("B=${B} failed:", e, open("gpurun_out/${TAG}_bench_b${B}.err").read()[-800:])
PY
done
