#!/bin/bash
# Runs on the GPU box: probes + kernel tests, each group in its own process, everything logged to gpurun_out/.
mkdir -p gpurun_out
LOG=gpurun_out/diag.txt
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv >> $LOG 2>&1
for combo in "0 0 0" "0 1 0" "1 0 0" "1 1 0" "0 0 1" "0 1 1" "1 1 1"; do
  timeout 300 python tools/gemm_probe.py $combo >> $LOG 2>&1 || echo "probe $combo exit $?" >> $LOG
done
for grp in "test_gemm" "test_rmsnorm or test_rope or test_swiglu or test_cast" "test_attention" "test_embed or test_ce or test_adamw" ; do
  echo "=== pytest -k '$grp'" >> $LOG
  timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "$grp" 2>&1 | tail -40 >> $LOG
done
tail -5 $LOG
