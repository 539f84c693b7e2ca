#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_api_round2_gpu.py -x -q -m gpu -k test_vit_bart_class_uses 2>&1 | tail -60 > gpurun_out/j25_tests.log
cat gpurun_out/j25_tests.log
