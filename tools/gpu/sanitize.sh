#!/bin/bash
# compute-sanitizer over the kernel tests at their small shapes (the reference has no race detection of its own;
# here every hand-written kernel gets memcheck, and the shared-memory-heavy SIMT kernels racecheck + synccheck).
# Not run in round 1. usage: gpurun --timeout 1500 -- 'bash tools/gpu/sanitize.sh r2s'
mkdir -p gpurun_out
TAG=${1:-r2s}
SMALL='not 8192 and not 4096 and not many_tiles and not llama7b and not full_geometry and not real_head_shape'
run() {  # $1 = tool, $2 = output suffix, $3 = pytest -k expression
  ( timeout 900 compute-sanitizer --tool "$1" --error-exitcode 9 --print-limit 20 \
      python -m pytest tests/test_kernels_gpu.py tests/test_jukebox_gpu.py -m gpu -q -x -p no:cacheprovider -k "$3" 2>&1 | tail -40 ) \
      > gpurun_out/${TAG}_$2.txt 2>&1
  echo "$1: $(grep -E 'passed|failed|ERROR SUMMARY' gpurun_out/${TAG}_$2.txt | tail -2 | tr '\n' ' ')"
}
run memcheck memcheck "$SMALL"
run racecheck racecheck "($SMALL) and (rmsnorm or swiglu or rope or ce_fwd or splice or adamw or colsum or pool or conv or sparse_attention)"
run synccheck synccheck "($SMALL) and (attention or gemm_bf16 and 128)"
