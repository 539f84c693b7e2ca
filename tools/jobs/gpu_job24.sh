#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/j24_tests.log
cat gpurun_out/j24_tests.log
timeout 300 python tools/attn_bench.py > gpurun_out/j24_attn_bench.txt 2>&1
cat gpurun_out/j24_attn_bench.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/j24_bench.json 2> gpurun_out/j24_bench.err
tail -c 1500 gpurun_out/j24_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/j24_bench.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l)
        print(d.get('metric'), d.get('value'), d.get('ms_per_step'), d.get('step_roofline',{}).get('frac'), d.get('e2e',{}).get('value'))
        if 'vit_bart' in d: print('vit_bart', d['vit_bart']['value'], d['vit_bart']['step_roofline']['frac'])
        if 'generate' in d: print('generate', {k:d['generate'].get(k) for k in ('value','ms_per_call')})
PY
