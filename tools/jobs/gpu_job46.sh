#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/j46_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j46_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j46_smoke.txt 2>&1
timeout 900 python bench.py > gpurun_out/j46_bench.json 2> gpurun_out/j46_bench.err
tail -3 gpurun_out/j46_pytest.log; tail -2 gpurun_out/j46_smoke.txt; tail -3 gpurun_out/j46_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/j46_bench.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l)
        print('train', d['value'], d['ms_per_step'], d['step_roofline']['frac'], 'e2e', d['e2e']['value'], 'u8', d['e2e_uint8_input']['value'], 'raw', d['e2e_raw_images'])
        print('gen', d['generate']['value'], 'vb', d['vit_bart']['value'], d['vit_bart']['step_roofline']['frac'], 'tr', d['transform']['value'], d['transform']['e2e']['value'])
PY
