#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/microbench_search.py 2>&1 | grep -v Warn | tail -8 > gpurun_out/j51_search.txt; cat gpurun_out/j51_search.txt
timeout 900 python -m pytest tests -q -m gpu -k "search or generate or beam or greedy or sample or golden or fused" 2>&1 | tail -4
