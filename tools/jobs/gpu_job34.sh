#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --no-generate --no-vit-bart > gpurun_out/j34_bench.json 2> gpurun_out/j34_bench.err
tail -5 gpurun_out/j34_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/j34_bench.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(json.dumps(d.get('transform'), indent=1))
PY
cat > /tmp/tr_one.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from mic_b200 import transforms
rng = np.random.RandomState(7)
imgs = [rng.randint(0, 256, (3, 375 + (i % 5) * 20, 500 + (i % 7) * 20)).astype(np.uint8) for i in range(256)]
bt = transforms.BatchTransform(224, 'cuda:0')
for _ in range(3): bt(imgs)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:resize_crop -s 2 -c 1 -o gpurun_out/j34_resize -f python /tmp/tr_one.py > gpurun_out/j34_ncu.log 2>&1
tail -2 gpurun_out/j34_ncu.log
