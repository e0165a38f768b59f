timeout 500 python -m pytest tests/test_reference_train_dropin_gpu.py tests/test_kernels_gpu.py tests/test_jukebox_gpu.py tests/test_model_gpu.py -m gpu -q --timeout 300 2>&1 | tail -12
timeout 200 python tools/bench_jukebox.py --precision bf16 --steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['clips_per_s'], d['ms_per_clip_by_stage'], {k:v['tflops'] for k,v in d['prior']['gemm_by_shape'].items()})"
timeout 200 python bench.py --no-extras --no-torch-baseline --no-cpu-baseline --steps 4 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['roofline']['achieved'])"
