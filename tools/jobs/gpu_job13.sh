#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/j13_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j13_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j13_smoke.txt 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/j13_bench.json 2> gpurun_out/j13_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/j13_bench_reference.json 2> gpurun_out/j13_bench_reference.err
ncu --kernel-name-base demangled --set full --clock-control none --import-source on -k regex:EpiCEStats -s 2 -c 1 -o gpurun_out/r02_cestats python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-generate --no-vit-bart > gpurun_out/j13_ncu_ce.log 2>&1
ncu --kernel-name-base demangled --set full --clock-control none --import-source on -k regex:EpiSearchPacked -s 5 -c 1 -o gpurun_out/r02_search python tools/microbench_search.py > gpurun_out/j13_ncu_search.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1700 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-generate --no-vit-bart > gpurun_out/j13_ncu_train.log 2>&1
tail -6 gpurun_out/j13_pytest.log; tail -2 gpurun_out/j13_smoke.txt; tail -3 gpurun_out/j13_bench.err; ls -la gpurun_out/r02_*.ncu-rep
