#!/bin/bash
# Evidence pass of round 2 (VERDICT r1 "Next round" item 9), ONE gpurun call on one GPU:
#   1. ncu launch list of one bench step (shares per kernel)                         -> <tag>_launches.csv
#   2. per-launch DRAM bytes + duration of every GEMM of the bench's own shapes      -> <tag>_gemm_dram.csv + <tag>_gemm_shapes.json
#   3. ncu --set full of the HBM-bound kernels and the attention kernels (real captures of adamw / sumsq / rmsnorm / swiglu /
#      ce / rope / attention fwd+bwd), of the GEMM kinds, and of the Jukebox kernels   -> <tag>_prof_*.ncu-rep
#   4. compute-sanitizer memcheck over the small-shape kernel tests                  -> <tag>_memcheck.txt
# usage: gpurun --timeout 2400 -- 'bash tools/gpu/r2_evidence.sh r2e'
mkdir -p gpurun_out
TAG=${1:-r2e}
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-torch-baseline --no-extras"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  $B > gpurun_out/${TAG}_ncu_bench.log 2>&1
echo "launch list: $(grep -c gpu__time_duration gpurun_out/${TAG}_launches.csv) rows"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
  -k regex:gemm --csv --log-file gpurun_out/${TAG}_gemm_dram.csv \
  python tools/gemm_shape_trace.py --out gpurun_out/${TAG}_gemm_shapes.json > gpurun_out/${TAG}_gemm_trace.log 2>&1
tail -1 gpurun_out/${TAG}_gemm_trace.log
S="python tools/gemm_shape_trace.py --out /tmp/shapes.json"
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"adamw_kernel|sumsq_partial|rmsnorm_fwd|rmsnorm_bwd|swiglu_fwd|ce_fwd|ce_bwd|rope_kernel|attn_fwd_tc|attn_bwd_dq_tc|attn_bwd_dkdv_tc|attn_bwd_prep|embed_splice" \
  -c 16 -o gpurun_out/${TAG}_prof_misc $S > gpurun_out/${TAG}_ncu_misc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 1 -c 10 \
  -o gpurun_out/${TAG}_prof_gemm $S > gpurun_out/${TAG}_ncu_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"resblock_kernel|conv_down_kernel|conv_k3_kernel|vq_encode|layernorm_kernel|jb_sparse_attn_tc|jb_pool|jb_embed" -c 14 \
  -o gpurun_out/${TAG}_prof_jukebox python tools/bench_jukebox.py --precision bf16 --depth 3 --steps 1 --warmup 0 \
  > gpurun_out/${TAG}_ncu_jukebox.log 2>&1
ls -la gpurun_out/${TAG}_prof_*.ncu-rep 2>/dev/null
SMALL='not 8192 and not 4096 and not 2048 and not 4160 and not many_tiles and not llama7b and not full_geometry and not real_head_shape'
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_kernels_gpu.py tests/test_jukebox_gpu.py tests/test_mpt_gpu.py -m gpu -q -x -p no:cacheprovider -k "$SMALL" 2>&1 | tail -40 ) \
    > gpurun_out/${TAG}_memcheck.txt 2>&1
echo "memcheck: $(grep -E 'passed|failed|ERROR SUMMARY' gpurun_out/${TAG}_memcheck.txt | tail -2 | tr '\n' ' ')"
