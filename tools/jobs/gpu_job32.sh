#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/j32_tests.log
cat gpurun_out/j32_tests.log
timeout 600 python tools/gemm_table.py 2>&1 | grep -v Warn > gpurun_out/j32_gemm_table.txt; tail -3 gpurun_out/j32_gemm_table.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/j32_bench.json 2> gpurun_out/j32_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/j32_bench.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l)
        print(d.get('metric'), d.get('value'), d.get('ms_per_step'), d.get('step_roofline',{}).get('frac'), d.get('e2e',{}).get('value'), d['roofline']['all_gemms']['frac'], d['clocks'])
        if 'vit_bart' in d: print('vit_bart', d['vit_bart']['value'], d['vit_bart']['step_roofline']['frac'])
        if 'generate' in d: print('generate', {k:d['generate'].get(k) for k in ('value','ms_per_call')})
PY
