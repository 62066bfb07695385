#!/bin/bash
# round 2, GPU call R: TMEM-drain micro-benchmark; shot chain with ff.net.2 folded into the [h | s1] stage (5 stages); shared-space hint
O=gpurun_out/r2aj; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -x > $O/tests.log 2>&1; echo "tests rc=$?"; tail -4 $O/tests.log | cut -c1-300
timeout 600 python bench.py --no-cpu --no-e2e --steps 20 --warmup 5 --dump-layers $O/layers.json > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cut -c1-600 $O/bench.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2aj/layers.json'))
for r in d['layers']:
    n=r['name'] if isinstance(r,dict) else r[0]; ms=r['ms'] if isinstance(r,dict) else r[1]
    if 'chain' in n or n in ('shot_time.block1.proj','shot_time.block2.proj','downs.0.0.block1.proj','downs.0.0.block2.proj','ups.3.0.block1.proj','mid_block1.block1.proj','ups.2.2.ff.net.0.0','init_conv','init_conv.pack','ups.3.0.block2.proj','final_res_block.block2.proj','downs.1.0.block2.proj'): print(f"{n:50s} {ms*1000:8.1f} us")
print('step', d['ms_per_step'])
PY
