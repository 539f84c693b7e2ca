#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/j40_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j40_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j40_smoke.txt 2>&1
timeout 900 python bench.py > gpurun_out/j40_bench.json 2> gpurun_out/j40_bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/j40_bench_ref.json 2> gpurun_out/j40_bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1700 --csv --log-file gpurun_out/j40_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-generate --no-vit-bart --no-transform > gpurun_out/j40_ncu.log 2>&1
timeout 600 python tools/gemm_table.py 2>&1 | grep -v Warn > gpurun_out/j40_gemm_table.txt
tail -3 gpurun_out/j40_pytest.log; tail -2 gpurun_out/j40_smoke.txt; tail -3 gpurun_out/j40_bench.err; tail -c 600 gpurun_out/j40_bench_ref.json; tail -2 gpurun_out/j40_gemm_table.txt
