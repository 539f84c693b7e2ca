#!/bin/bash
mkdir -p gpurun_out
for t in memcheck racecheck synccheck; do
  timeout 400 compute-sanitizer --tool $t python tools/sanitize_run.py attn > gpurun_out/j43_$t.txt 2>&1
  echo "== $t rc=$?"; tail -4 gpurun_out/j43_$t.txt
done
timeout 400 compute-sanitizer --tool memcheck python tools/sanitize_run.py train > gpurun_out/j43_memcheck_train.txt 2>&1
echo "== memcheck train rc=$?"; tail -3 gpurun_out/j43_memcheck_train.txt
