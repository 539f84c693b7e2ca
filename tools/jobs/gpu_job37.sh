#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $1"; env $1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/dp_tail_sweep.py 2 2>&1 | grep -E "^rep|Error|error"; }
{ run "NCCL_DEBUG=WARN"; run "NCCL_MAX_CTAS=4"; run "NCCL_MAX_CTAS=16"; run "NCCL_ALGO=Ring"; } > gpurun_out/j37_nccl_sweep.txt 2>&1
cat gpurun_out/j37_nccl_sweep.txt
