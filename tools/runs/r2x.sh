#!/bin/bash
# round 2, GPU call X: chain kernels with the GELU outputs handed to the next GEMM through tensor memory (NDIFF_CHAIN_TS=1): tests + A/B
O=gpurun_out/r2x; mkdir -p $O
NDIFF_CHAIN_TS=1 timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_net.py -m gpu -q -x -k "chain or net or forward or teacher or golden or step" > $O/tests_ts.log 2>&1; echo "tests(ts) rc=$?"; tail -5 $O/tests_ts.log | cut -c1-400
for t in 0 1 0 1; do
NDIFF_CHAIN_TS=$t timeout 600 python bench.py --no-cpu --no-e2e --steps 20 --warmup 5 --dump-layers $O/layers_ts$t.json > $O/bench_ts$t.json 2> $O/bench_ts$t.err; echo "bench ts=$t rc=$?"
python - <<PY
import json
d=json.load(open('$O/layers_ts$t.json'))
print('ts=$t step', round(d['ms_per_step'],3), {(r['name'] if isinstance(r,dict) else r[0])[:28]: round((r['ms'] if isinstance(r,dict) else r[1])*1000,1) for r in d['layers'] if 'chain' in (r['name'] if isinstance(r,dict) else r[0])})
PY
done
