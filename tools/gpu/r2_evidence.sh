#!/bin/bash
# Evidence pass of round 2 (VERDICT r1 "Next round" item 9), ONE gpurun call on one GPU. Reports are summarised ON THE BOX
# (tools/summarize_ncu.py) and the .ncu-rep files deleted: gpurun merges at most 64 MiB back.
#   1. ncu launch list of one bench step (shares per kernel)                         -> <tag>_launches.csv
#   2. per-launch DRAM bytes + duration of every GEMM of the bench's own shapes      -> <tag>_gemm_dram.csv + <tag>_gemm_shapes.json
#   3. ncu --set full of the HBM-bound kernels and the attention kernels, of the GEMM kinds, of the Jukebox kernels and of
#      the prior's four GEMMs (with source-level stall samples)                        -> <tag>_*_full.md / _hbm.md / _source.csv
#   4. (optional, SAN=1) compute-sanitizer memcheck over the small-shape kernel tests -> <tag>_memcheck.txt
# usage: gpurun --timeout 2400 -- 'bash tools/gpu/r2_evidence.sh r2f'
mkdir -p gpurun_out
TAG=${1:-r2f}
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-torch-baseline --no-extras"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  $B > gpurun_out/${TAG}_ncu_bench.log 2>&1
echo "launch list: $(grep -c gpu__time_duration gpurun_out/${TAG}_launches.csv) rows"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
  -k regex:gemm --csv --log-file gpurun_out/${TAG}_gemm_dram.csv \
  python tools/gemm_shape_trace.py --out gpurun_out/${TAG}_gemm_shapes.json > gpurun_out/${TAG}_gemm_trace.log 2>&1
tail -1 gpurun_out/${TAG}_gemm_trace.log
S="python tools/gemm_shape_trace.py --out /tmp/shapes.json"
summ() {  # $1 = report stem; writes <stem>_full.md and <stem>_hbm.md, removes the report
  python tools/summarize_ncu.py full gpurun_out/$1.ncu-rep gpurun_out/$1_full.md > /dev/null 2>&1
  python tools/summarize_ncu.py hbm gpurun_out/$1.ncu-rep gpurun_out/$1_hbm.md > /dev/null 2>&1
  rm -f gpurun_out/$1.ncu-rep
}
timeout 900 ncu --set full --clock-control none \
  -k regex:"adamw_kernel|sumsq_partial|rmsnorm_fwd|rmsnorm_bwd|swiglu_fwd|ce_fwd|ce_bwd|rope_kernel|attn_fwd_tc|attn_bwd_dq_tc|attn_bwd_dkdv_tc|attn_bwd_prep|embed_splice" \
  -c 16 -o gpurun_out/${TAG}_prof_misc $S > gpurun_out/${TAG}_ncu_misc.log 2>&1
summ ${TAG}_prof_misc
timeout 900 ncu --set full --clock-control none -k regex:gemm2_kernel -s 1 -c 10 \
  -o gpurun_out/${TAG}_prof_gemm $S > gpurun_out/${TAG}_ncu_gemm.log 2>&1
summ ${TAG}_prof_gemm
timeout 900 ncu --set full --clock-control none \
  -k regex:"resblock_kernel|conv_down_kernel|conv_k3_kernel|vq_encode|layernorm_kernel|jb_sparse_attn_tc|jb_pool|jb_embed" -c 14 \
  -o gpurun_out/${TAG}_prof_jukebox python tools/bench_jukebox.py --precision bf16 --depth 3 --steps 1 --warmup 0 \
  > gpurun_out/${TAG}_ncu_jukebox.log 2>&1
summ ${TAG}_prof_jukebox
# the prior's four GEMMs of block 0 (c_attn fp32 out, c_proj fp32 out + fp32 residual, fc bf16 quick-GELU, fc_proj) with
# source-level stall samples: where does the fp32 epilogue wait?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -c 4 \
  -o gpurun_out/${TAG}_prof_prior_gemm python tools/bench_jukebox.py --precision bf16 --depth 1 --steps 1 --warmup 0 \
  > gpurun_out/${TAG}_ncu_prior_gemm.log 2>&1
ncu -i gpurun_out/${TAG}_prof_prior_gemm.ncu-rep --page source --csv > gpurun_out/${TAG}_prior_gemm_source_all.csv 2>/dev/null
python - <<PY
import csv, io
rows = list(csv.reader(open("gpurun_out/${TAG}_prior_gemm_source_all.csv", errors="replace")))
hdr = next((r for r in rows if "Source" in r or "# Samples" in " ".join(r)), None)
print("source page columns:", hdr[:30] if hdr else None, "rows", len(rows))
PY
head -c 3000000 gpurun_out/${TAG}_prior_gemm_source_all.csv > gpurun_out/${TAG}_prior_gemm_source.csv
rm -f gpurun_out/${TAG}_prior_gemm_source_all.csv
summ ${TAG}_prof_prior_gemm
if [ "${SAN:-0}" = "1" ]; then
SMALL='not 8192 and not 4096 and not 2048 and not 4160 and not many_tiles and not llama7b and not full_geometry and not real_head_shape'
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_kernels_gpu.py tests/test_jukebox_gpu.py tests/test_mpt_gpu.py -m gpu -q -x -p no:cacheprovider -k "$SMALL" 2>&1 | tail -40 ) \
    > gpurun_out/${TAG}_memcheck.txt 2>&1
echo "memcheck: $(grep -E 'passed|failed|ERROR SUMMARY' gpurun_out/${TAG}_memcheck.txt | tail -2 | tr '\n' ' ')"
fi
du -sh gpurun_out; ls -la gpurun_out
