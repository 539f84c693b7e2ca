#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/j4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j4_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j4_smoke.txt 2>&1
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/j4_bench.json 2> gpurun_out/j4_bench.err
tail -40 gpurun_out/j4_pytest.log
tail -3 gpurun_out/j4_smoke.txt
tail -3 gpurun_out/j4_bench.err
