#!/bin/bash
# same-box A/B: the tree of commit fe19805 (ab_base/) against the working tree, alternating
mkdir -p gpurun_out
for i in 1 2; do
  (cd ab_base && timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-generate --no-vit-bart 2>/dev/null | tail -1 > ../gpurun_out/j39_base_$i.json)
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-generate --no-vit-bart 2>/dev/null | tail -1 > gpurun_out/j39_new_$i.json
done
python - <<'PY'
import json
for n in ("base_1","new_1","base_2","new_2"):
    d=json.loads(open(f"gpurun_out/j39_{n}.json").read())
    print(n, round(d['value'],1), round(d['ms_per_step'],2), round(d['step_roofline']['frac'],4), round(d['roofline']['all_gemms']['frac'],4), d['clocks']['sm_mhz'])
PY
