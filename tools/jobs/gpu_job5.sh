#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/j5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j5_pytest.log
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-generate > gpurun_out/j5_bench.json 2> gpurun_out/j5_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1700 --csv --log-file gpurun_out/j5_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-generate > gpurun_out/j5_ncu_bench.log 2>&1
tail -15 gpurun_out/j5_pytest.log
tail -3 gpurun_out/j5_bench.err
