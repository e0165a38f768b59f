#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r1d}
echo "=== pytest -m gpu" > gpurun_out/${TAG}_pytest.txt
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -40 >> gpurun_out/${TAG}_pytest.txt
tail -6 gpurun_out/${TAG}_pytest.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1; tail -1 gpurun_out/${TAG}_smoke.txt
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 1200 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 1500 python bench.py --steps 5 --warmup 3 --batch 16 --no-cpu-baseline > gpurun_out/${TAG}_bench_b16.json 2> gpurun_out/${TAG}_bench_b16.err
tail -c 600 gpurun_out/${TAG}_bench_b16.json
