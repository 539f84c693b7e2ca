#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_transform_gpu.py -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py --no-cpu-baseline --no-generate --no-vit-bart > gpurun_out/j47_bench.json 2> gpurun_out/j47_bench.err
tail -3 gpurun_out/j47_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/j47_bench.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l)
        print('train', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'u8', d['e2e_uint8_input']['value'], 'raw', d['e2e_raw_images']['value'], d['e2e_raw_images']['ms_per_step'])
        print('tr', d['transform']['value'], d['transform']['e2e'])
PY
