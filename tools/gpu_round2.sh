#!/bin/bash
# GPU round: 2-CTA GEMM validation + timing, full test suite, smoke, bench in both GEMM modes.
mkdir -p gpurun_out
TAG=${1:-r1b}
LOG=gpurun_out/${TAG}_probe.txt
: > $LOG
for mode in 1 0; do
  for combo in "0 0 0" "0 1 0" "1 1 0"; do
    LLARK_GEMM_2CTA=$mode timeout 300 python tools/gemm_probe.py $combo >> $LOG 2>&1 || echo "probe mode=$mode $combo exit $?" >> $LOG
  done
done
LLARK_GEMM_2CTA=1 timeout 300 python tools/gemm_probe.py 0 0 1 >> $LOG 2>&1
grep -E "timing|EXC|exit|rel=[1-9]\.[0-9]*e[+-]0[0-1]" $LOG | tail -40
echo "=== pytest -m gpu" > gpurun_out/${TAG}_pytest.txt
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -60 >> gpurun_out/${TAG}_pytest.txt
tail -15 gpurun_out/${TAG}_pytest.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1; tail -2 gpurun_out/${TAG}_smoke.txt
LLARK_GEMM_2CTA=1 timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_2cta.json 2> gpurun_out/${TAG}_bench_2cta.err
tail -c 2500 gpurun_out/${TAG}_bench_2cta.json; tail -3 gpurun_out/${TAG}_bench_2cta.err
LLARK_GEMM_2CTA=0 timeout 1500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_1cta.json 2> gpurun_out/${TAG}_bench_1cta.err
tail -c 1500 gpurun_out/${TAG}_bench_1cta.json
