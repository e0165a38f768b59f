mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6 > gpurun_out/r2g_pytest.txt; cat gpurun_out/r2g_pytest.txt
cp gpurun_out/parity_measured.txt gpurun_out/r2g_parity_measured.txt 2>/dev/null
for v in "PF1:" "PF0:LLARK_GEMM_PREFETCH=0"; do n=${v%%:*}; e=${v#*:}; env $e timeout 300 python tools/bench_jukebox.py --precision bf16 --steps 2 > gpurun_out/r2g_jukebox_$n.json 2>/dev/null; python - <<PY
import json
d=json.loads(open("gpurun_out/r2g_jukebox_$n.json").read().strip().splitlines()[-1])
print("$n", round(d["clips_per_s"],2), d["ms_per_clip_by_stage"]["prior"], {k:v["tflops"] for k,v in d["prior"]["gemm_by_shape"].items()})
PY
done
X="--no-extras --no-torch-baseline --no-cpu-baseline --steps 4 --warmup 3"
for v in "default:" "noside:LLARK_WGRAD_STREAM=0" "nopf:LLARK_GEMM_PREFETCH=0" "l2_48:LLARK_GEMM_L2_MB=48" "l2_32:LLARK_GEMM_L2_MB=32" "default2:"; do n=${v%%:*}; e=${v#*:}; env $e timeout 300 python bench.py $X > gpurun_out/r2g_bench_$n.json 2>/dev/null; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2g_bench_$n.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("$n", "value", round(d["value"],2), "ms", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"],2), "gemm", round(r["achieved"],1), "pass_ms", round(r.get("pass_ms_per_step",0),1))
    if "$n" in ("default","l2_48","nopf"):
        for k,v in r["gemm_by_shape"].items(): print("    ",k,v["tflops"])
except Exception as ex: print("$n failed", ex)
PY
done
