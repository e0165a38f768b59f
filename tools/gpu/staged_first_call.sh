#!/bin/bash
# First gpurun call of the next round: everything that was written after the round-1 GPU budget was spent.
#   1. the staged parity tests (KV-cache decode attention + generate with cache; MPT LayerNorm / GELU / ALiBi attention
#      kernels and the whole MPT engine step against the pinned oracle; attention at S = 2048 / 4096 / 4160; activation
#      recompute bit-identity; the reference's 4-step training trajectory)           -> <tag>_staged_pytest.txt
#   2. compute-sanitizer memcheck over the staged per-op tests (small shapes only)  -> <tag>_staged_memcheck.txt
#   3. the configs[4] line (CLAP + MPT-1B training step)                            -> <tag>_bench_mpt.json
#   4. the regular suite + bench, so the numbers come from the same box            -> <tag>_pytest.txt, <tag>_bench.json
# usage: gpurun --timeout 1500 -- 'bash tools/gpu/staged_first_call.sh r2a'
mkdir -p gpurun_out
TAG=${1:-r2a}
( LLARK_STAGED=1 timeout 600 python -m pytest tests/test_mpt_gpu.py tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q \
    -k "mpt or layernorm_forward or exact_gelu or alibi or decode_step or kv_cache or long_sequences or recompute or trajectory or very_few_rows or gemv or frozen_backbone" 2>&1 | tail -40 ) > gpurun_out/${TAG}_staged_pytest.txt 2>&1
tail -3 gpurun_out/${TAG}_staged_pytest.txt
( LLARK_STAGED=1 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_mpt_gpu.py -m gpu -q -x \
    -k "layernorm_forward or exact_gelu and not 2457600 or alibi and 128" 2>&1 | tail -30 ) > gpurun_out/${TAG}_staged_memcheck.txt 2>&1
tail -3 gpurun_out/${TAG}_staged_memcheck.txt
timeout 600 python tools/bench_mpt.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_mpt.json 2> gpurun_out/${TAG}_bench_mpt.err
tail -1 gpurun_out/${TAG}_bench_mpt.json | cut -c1-600
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > gpurun_out/${TAG}_pytest.txt 2>&1; tail -2 gpurun_out/${TAG}_pytest.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -1 gpurun_out/${TAG}_bench.json | cut -c1-400
