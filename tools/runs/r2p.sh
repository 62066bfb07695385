#!/bin/bash
O=gpurun_out/r2p; mkdir -p $O
NDIFF_CHAIN2=1 timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_net.py -m gpu -q -x -k "chain or forward_matches or teacher" > $O/tests_chain2.log 2>&1; echo "chain2 tests rc=$?"; tail -4 $O/tests_chain2.log | cut -c1-300
timeout 300 python bench.py --no-e2e --no-cpu --steps 20 --warmup 5 --dump-layers $O/layers_chain1.json > $O/bench_chain1.json 2> $O/bench.err; echo "bench1 rc=$?"
NDIFF_CHAIN2=1 timeout 300 python bench.py --no-e2e --no-cpu --steps 20 --warmup 5 --dump-layers $O/layers_chain2.json > $O/bench_chain2.json 2>> $O/bench.err; echo "bench2 rc=$?"
python - <<'PY'
import json
a=json.load(open('gpurun_out/r2p/layers_chain1.json')); b=json.load(open('gpurun_out/r2p/layers_chain2.json'))
da={r[0]:r[1] for r in a['layers']}; db={r[0]:r[1] for r in b['layers']}
for k in da:
    if 'chain' in k: print(f"{k:50s} {da[k]*1e3:8.1f} -> {db.get(k,0)*1e3:8.1f}")
print(a['ms_per_step'], b['ms_per_step'])
PY
