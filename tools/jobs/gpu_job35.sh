#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_transform_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 2 --warmup 3 --no-generate --no-vit-bart --no-cpu-baseline > gpurun_out/j35_bench.json 2> gpurun_out/j35_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/j35_bench.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l)['transform']; print(d['value'], d['ms_per_batch'], d['roofline']['frac'], d['e2e'])
PY
