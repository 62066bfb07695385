#!/bin/bash
O=gpurun_out/r2n; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_net.py tests/test_gpu_dim48.py -m gpu -q -x -k "chain or forward or teacher or dim48" > $O/tests.log 2>&1; echo "tests rc=$?"; tail -4 $O/tests.log | cut -c1-300
timeout 300 python bench.py --no-e2e --no-cpu --steps 20 --warmup 5 --dump-layers $O/layers.json > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cut -c1-330 $O/bench.json
python - <<'PY'
import json
z=json.load(open('gpurun_out/r2n/layers.json'))
for n,t,f,b in z['layers']:
    if 'chain' in n: print(f"{n:50s} {t*1e3:8.1f} us")
print(z['ms_per_step'], sum(r[1] for r in z['layers']))
PY
