#!/bin/bash
set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/j16_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j16_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j16_smoke.txt 2>&1
timeout 900 python bench.py > gpurun_out/j16_bench.json 2> gpurun_out/j16_bench.err
tail -4 gpurun_out/j16_pytest.log; tail -2 gpurun_out/j16_smoke.txt; tail -3 gpurun_out/j16_bench.err
