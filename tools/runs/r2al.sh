#!/bin/bash
# round 2, GPU call AL: staged TMA store in the epilogue of the N = 64 conv kernels: op tests, same-box A/B, whole suite
O=gpurun_out/r2al; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "conv or downsample or upsample or groupnorm" > $O/tests_ops.log 2>&1; echo "op tests rc=$?"; tail -3 $O/tests_ops.log | cut -c1-300
for t in 1 0 1 0; do
NDIFF_NO_STAGED_STORE=$t timeout 600 python bench.py --no-cpu --no-e2e --steps 20 --warmup 5 --dump-layers $O/layers_$t.json > $O/bench_$t.json 2> $O/bench_$t.err; echo "bench no_staged=$t rc=$?"; tail -2 $O/bench_$t.err | cut -c1-300
python - <<PY
import json
d=json.load(open('$O/layers_$t.json'))
g=lambda r:(r['name'],r['ms']) if isinstance(r,dict) else (r[0],r[1])
want=('shot_time.block1.proj','shot_time.block2.proj','downs.0.0.block1.proj','downs.0.0.block2.proj','ups.3.0.block1.proj','mid_block1.block1.proj','init_conv','ups.2.3.1','downs.0.3.1','ups.3.3')
print('no_staged=$t step', round(d['ms_per_step'],3), {n[:24]: round(ms*1000,1) for n,ms in map(g,d['layers']) if n in want})
PY
done
timeout 1800 python -m pytest tests -m gpu -q -x > $O/tests.log 2>&1; echo "tests rc=$?"; tail -4 $O/tests.log | cut -c1-300
