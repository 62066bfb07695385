#!/bin/bash
# round 2, GPU call W: A/B of the chain kernels with one vs two threads per pixel (same box), after the issue / table / walk fixes
O=gpurun_out/r2w; mkdir -p $O
NDIFF_CHAIN_TPP=2 timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_net.py -m gpu -q -x > $O/tests_tpp2.log 2>&1; echo "tests(tpp2) rc=$?"; tail -3 $O/tests_tpp2.log | cut -c1-300
for t in 1 2 1 2; do
NDIFF_CHAIN_TPP=$t timeout 600 python bench.py --no-cpu --no-e2e --steps 20 --warmup 5 --dump-layers $O/layers_tpp$t.json > $O/bench_tpp$t.json 2> $O/bench_tpp$t.err; echo "bench tpp=$t rc=$?"
python - <<PY
import json
d=json.load(open('$O/layers_tpp$t.json'))
print('tpp=$t step', round(d['ms_per_step'],3), {(r['name'] if isinstance(r,dict) else r[0])[:28]: round((r['ms'] if isinstance(r,dict) else r[1])*1000,1) for r in d['layers'] if 'chain' in (r['name'] if isinstance(r,dict) else r[0])})
PY
done
