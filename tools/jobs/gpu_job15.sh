#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_api_round2_gpu.py tests/test_generate_edge_gpu.py tests/test_beam_kernels_gpu.py tests/test_fused_decoder_gpu.py -q > gpurun_out/j15_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j15_pytest.log
timeout 200 python tools/microbench_search.py > gpurun_out/j15_search.txt 2>&1
ncu --kernel-name-base demangled --set full --clock-control none --import-source on -k regex:EpiSearchPacked -s 5 -c 1 -o gpurun_out/r02_search python tools/microbench_search.py > gpurun_out/j15_ncu_search.log 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-vit-bart > gpurun_out/j15_bench.json 2> gpurun_out/j15_bench.err
tail -4 gpurun_out/j15_pytest.log; cat gpurun_out/j15_search.txt; tail -2 gpurun_out/j15_bench.err
