#!/bin/bash
# round-2 GPU job 2: tests incl. full-size goldens with the new decoder defaults, stress, bench (new bench line), profiles
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/j2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j2_pytest.log
timeout 200 python tools/stress_fused.py > gpurun_out/j2_stress.txt 2>&1
timeout 200 python tools/profile_decoder_step.py 40 64 0 > gpurun_out/j2_phases.txt 2>&1
timeout 200 python tools/profile_decoder_step.py 40 64 1 > gpurun_out/j2_phases_writerfence.txt 2>&1
timeout 600 python bench.py --steps 6 --warmup 3 > gpurun_out/j2_bench.json 2> gpurun_out/j2_bench.err
tail -5 gpurun_out/j2_pytest.log
tail -3 gpurun_out/j2_bench.err
