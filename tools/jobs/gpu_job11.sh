#!/bin/bash
# round-2 evidence job: tests, bench, launch lists, ncu --set full of the dominant kernels, sanitizers
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/j11_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j11_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/j11_bench.json 2> gpurun_out/j11_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1700 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-generate --no-vit-bart > gpurun_out/j11_ncu_train.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_gen_launches.csv python tools/profile_generate.py 64 64 > gpurun_out/j11_ncu_gen.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:EpiCEStats -s 2 -c 1 -o gpurun_out/r02_cestats python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-generate --no-vit-bart > gpurun_out/j11_ncu_ce.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:decoder_step_kernel -s 30 -c 1 -o gpurun_out/r02_decoder_step python tools/profile_decoder_step.py 40 > gpurun_out/j11_ncu_dec.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:EpiSearchPacked -s 5 -c 1 -o gpurun_out/r02_search python tools/microbench_search.py > gpurun_out/j11_ncu_search.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:act_bwd_colsum -s 40 -c 12 -o gpurun_out/r02_actbwd python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-generate --no-vit-bart > gpurun_out/j11_ncu_act.log 2>&1
for pos in 10 40 62; do timeout 200 python tools/profile_decoder_step.py $pos >> gpurun_out/r02_decoder_step_phases.txt 2>&1; done
timeout 120 python tools/microbench_barrier.py > gpurun_out/r02_barrier_microbench.txt 2>&1
for t in memcheck racecheck synccheck; do
  timeout 420 compute-sanitizer --tool $t --print-limit 20 python tools/sanitize_run.py > gpurun_out/r02_sanitizer_$t.txt 2>&1
  echo "rc=$?" >> gpurun_out/r02_sanitizer_$t.txt
done
timeout 300 python tools/determinism_check.py > gpurun_out/r02_determinism.txt 2>&1
tail -12 gpurun_out/j11_pytest.log
tail -3 gpurun_out/j11_bench.err
ls -la gpurun_out | grep r02_
