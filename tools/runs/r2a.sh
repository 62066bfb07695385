#!/bin/bash
# round 2, GPU call A: the tcgen05.mma.ws experiment, a baseline layer table, and the new full-shape parity tests
O=gpurun_out/r2a; mkdir -p $O
NDIFF_TEST_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_gpu_ops.py -m gpu -k experimental -x -q > $O/ws_test.log 2>&1; echo "ws_test rc=$?"
python bench.py --no-e2e --no-cpu --steps 20 --warmup 5 --dump-layers $O/base_layers.json > $O/base.json 2> $O/base.err
NDIFF_EXPERIMENT_WS=1 timeout 200 python bench.py --no-e2e --no-cpu --steps 20 --warmup 5 --dump-layers $O/ws1_layers.json > $O/ws1.json 2> $O/ws1.err
NDIFF_EXPERIMENT_WS=2 timeout 200 python bench.py --no-e2e --no-cpu --steps 20 --warmup 5 --dump-layers $O/ws2_layers.json > $O/ws2.json 2> $O/ws2.err
tail -3 $O/ws_test.log
cat $O/base.json $O/ws1.json $O/ws2.json | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_fullshape.py -m gpu -x -q -s > $O/fullshape.log 2>&1; echo "fullshape rc=$?"
tail -40 $O/fullshape.log
timeout 600 python -m pytest tests/test_gpu_dim48.py -m gpu -q -s > $O/dim48.log 2>&1; echo "dim48 rc=$?"
tail -60 $O/dim48.log
