#!/bin/bash
# round 2, GPU call B: whole GPU suite (regression after the ws default / dim 48 / ABI changes) + a full default bench line
O=gpurun_out/r2b; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?"; tail -5 $O/tests.log
timeout 600 python bench.py --dump-layers $O/layers.json > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
cut -c1-1500 $O/bench.json
timeout 300 python bench.py --impl reference --steps 6 --warmup 2 > $O/bench_ref.json 2>> $O/bench.err; cut -c1-300 $O/bench_ref.json
timeout 300 python bench.py --dim 48 --patch 512 --batch 16 --no-cpu --steps 10 --dump-layers $O/layers_dim48.json > $O/bench_dim48.json 2>> $O/bench.err; cut -c1-400 $O/bench_dim48.json
