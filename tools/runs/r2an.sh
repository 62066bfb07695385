#!/bin/bash
# round 2, GPU call AN: micro-batch size under the power cap (smaller tensors stay in L2 longer; more launches)
O=gpurun_out/r2an; mkdir -p $O
for mb in 64 32 16 64; do
timeout 600 python bench.py --no-cpu --no-e2e --steps 20 --warmup 5 --micro-batch $mb > $O/bench_mb$mb.json 2> $O/bench_mb$mb.err; echo "mb=$mb rc=$?"
python - <<PY
import json
d=json.loads(open('$O/bench_mb$mb.json').read().strip().splitlines()[-1]); print('mb=$mb', d['value'], d['ms_per_step'], d['clocks'])
PY
done
