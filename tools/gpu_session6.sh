#!/bin/bash
# one GPU call: full GPU suite, frame bench (BASELINE configs[3] at N=1), bench.py
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/s6b_tests.log 2>&1; echo "suite rc=$?"
tail -3 gpurun_out/s6b_tests.log
timeout 150 python tools/bench_frame.py --out gpurun_out/s6b_frame.json > gpurun_out/s6b_frame.log 2>&1; echo "frame rc=$?"
tail -2 gpurun_out/s6b_frame.log | cut -c1-900
timeout 200 python bench.py > gpurun_out/s6b_bench.json 2> gpurun_out/s6b_bench.err; echo "bench rc=$?"
cut -c1-1500 gpurun_out/s6b_bench.json
