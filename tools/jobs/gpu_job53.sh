#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/j53_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j53_pytest.log
tail -3 gpurun_out/j53_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j53_smoke.txt 2>&1; tail -1 gpurun_out/j53_smoke.txt
timeout 600 python tools/gemm_table.py 2>&1 | grep -v Warn > gpurun_out/j53_gemm_table.txt; tail -1 gpurun_out/j53_gemm_table.txt
BNS=0 timeout 300 python tools/gemm_epilogue_bench.py 2>&1 | grep -v Warn > gpurun_out/j53_gemm_epi.txt; cat gpurun_out/j53_gemm_epi.txt
timeout 900 python bench.py > gpurun_out/j53_bench.json 2> gpurun_out/j53_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/j53_bench.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l)
        print('train', d['value'], d['ms_per_step'], d['step_roofline']['frac'], d['roofline']['all_gemms']['frac'], d['roofline']['frac'], d['clocks']['sm_mhz'], 'e2e', d['e2e']['value'], d['e2e_uint8_input']['value'], d['e2e_raw_images']['value'])
        g=d['generate']; print('gen', g['value'], g['ms_per_call'], g['roofline']['frac'], g['e2e']['value'], g['lm_head_search_kernel']['ms_per_launch'], g['decoder_step_kernel']['ms_per_launch'])
        print('vb', d['vit_bart']['value'], d['vit_bart']['ms_per_step'], d['vit_bart']['step_roofline']['frac'], 'tr', d['transform']['value'], d['transform']['e2e']['value'])
PY
