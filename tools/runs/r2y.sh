#!/bin/bash
# round 2, GPU call Y: tensor-memory operands everywhere in the chain kernels (default): whole GPU suite + A/B against NDIFF_CHAIN_TS=0
O=gpurun_out/r2y; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -x > $O/tests.log 2>&1; echo "tests rc=$?"; tail -4 $O/tests.log | cut -c1-300
for t in 0 1 0 1; do
NDIFF_CHAIN_TS=$t timeout 600 python bench.py --no-cpu --no-e2e --steps 20 --warmup 5 --dump-layers $O/layers_ts$t.json > $O/bench_ts$t.json 2> $O/bench_ts$t.err; echo "bench ts=$t rc=$?"
python - <<PY
import json
d=json.load(open('$O/layers_ts$t.json'))
print('ts=$t step', round(d['ms_per_step'],3), {(r['name'] if isinstance(r,dict) else r[0])[:28]: round((r['ms'] if isinstance(r,dict) else r[1])*1000,1) for r in d['layers'] if 'chain' in (r['name'] if isinstance(r,dict) else r[0])})
PY
done
