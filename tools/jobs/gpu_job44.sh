#!/bin/bash
# same-box A/B: activation backward fused into the fc2 dgrad GEMM epilogue (MIC_FUSE_ACT_BWD=1) vs the separate pass
mkdir -p gpurun_out
for i in 1 2; do
  for f in 0 1; do
    MIC_FUSE_ACT_BWD=$f timeout 600 python bench.py --steps 15 --warmup 3 --no-cpu-baseline --no-generate --no-vit-bart --no-transform 2>/dev/null | tail -1 > gpurun_out/j44_fuse${f}_$i.json
  done
done
python - <<'PY'
import json
for n in ("fuse0_1","fuse1_1","fuse0_2","fuse1_2"):
    d=json.loads(open(f"gpurun_out/j44_{n}.json").read())
    print(n, round(d['value'],1), round(d['ms_per_step'],2), round(d['step_roofline']['frac'],4), round(d['roofline']['all_gemms']['frac'],4), d['clocks']['sm_mhz'], d['config']['loss_last'])
PY
