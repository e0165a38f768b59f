#!/bin/bash
# One GPU round: tests, bench, ncu launch list + one full capture of the GEMM core. Logs under gpurun_out/.
mkdir -p gpurun_out
TAG=${1:-r1}
echo "=== pytest -m gpu" > gpurun_out/${TAG}_pytest.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 >> gpurun_out/${TAG}_pytest.txt
tail -3 gpurun_out/${TAG}_pytest.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1; tail -2 gpurun_out/${TAG}_smoke.txt
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 3000 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
if [ "$2" != "noncu" ]; then
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 40 -c 3 -o gpurun_out/${TAG}_prof_gemm \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/ | tail -12
fi
