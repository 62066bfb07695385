#!/bin/bash
O=gpurun_out/r2m; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_dim48.py tests/test_gpu_fullshape.py -m gpu -q -x -s -k "not free_running" > $O/tests.log 2>&1; echo "tests rc=$?"; grep -E "downs.2.2|downs.3.2|ups.0.2|ups.1.2|ups.2.2|passed|failed" $O/tests.log | head -30
timeout 300 python bench.py --no-e2e --no-cpu --steps 20 --warmup 5 --dump-layers $O/layers.json > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cut -c1-330 $O/bench.json
NDIFF_NO_WS=0 python - <<'PY'
import json
z=json.load(open('gpurun_out/r2m/layers.json'))
for n,t,f,b in z['layers']:
    if '.2.' in n or n.endswith('.norm2'):
        print(f"{n:40s} {t*1e3:8.1f} us")
print(z['ms_per_step'], sum(r[1] for r in z['layers']))
PY
