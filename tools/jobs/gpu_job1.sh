#!/bin/bash
# round-2 GPU job 1: regression tests, sanitizer evidence, decoder-step tuning switches, barrier microbenchmark
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/j1_smi.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/j1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j1_pytest.log
timeout 120 python tools/microbench_barrier.py > gpurun_out/j1_barrier.txt 2>&1
for o in 0 1 2 3; do
  timeout 200 python tools/profile_decoder_step.py 40 64 $o > gpurun_out/j1_phases_opts$o.txt 2>&1
done
for t in memcheck racecheck synccheck; do
  timeout 420 compute-sanitizer --tool $t --print-limit 20 python tools/sanitize_run.py > gpurun_out/j1_sanitizer_$t.txt 2>&1
  echo "rc=$?" >> gpurun_out/j1_sanitizer_$t.txt
done
for o in 0 3; do
  MIC_DECODER_OPTS=$o timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/j1_bench_opts$o.json 2> gpurun_out/j1_bench_opts$o.err
done
tail -3 gpurun_out/j1_pytest.log
