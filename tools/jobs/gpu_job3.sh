#!/bin/bash
set -x
for o in 0 1; do
  MIC_DECODER_OPTS=$o timeout 300 python tools/determinism_check.py > gpurun_out/j3_determinism_opts$o.txt 2>&1
  MIC_DECODER_OPTS=$o timeout 200 python tools/stress_fused.py > gpurun_out/j3_stress_opts$o.txt 2>&1
done
tail -9 gpurun_out/j3_determinism_opts*.txt; tail -2 gpurun_out/j3_stress_opts*.txt
