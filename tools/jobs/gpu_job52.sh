#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/j52_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j52_pytest.log
tail -3 gpurun_out/j52_pytest.log
timeout 600 python tools/gemm_table.py 2>&1 | grep -v Warn > gpurun_out/j52_gemm_table.txt; grep -E " 4096    1024  0  1|3072     768  0  1|total" gpurun_out/j52_gemm_table.txt
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/j52_bench.json 2> gpurun_out/j52_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/j52_bench.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l)
        print('train', d['value'], d['ms_per_step'], d['step_roofline']['frac'], d['roofline']['all_gemms']['frac'], d['clocks']['sm_mhz'])
        g=d['generate']; print('gen', g['value'], g['ms_per_call'], g['roofline']['frac'], g['lm_head_search_kernel'])
        print('vb', d['vit_bart']['value'], d['vit_bart']['step_roofline']['frac'])
PY
