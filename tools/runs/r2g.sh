#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -s -x > $O/train_tests.log 2>&1; echo "train tests rc=$?"; tail -12 $O/train_tests.log | cut -c1-300
timeout 600 python tools/bench_train.py --steps 10 --warmup 3 --out $O/train_n1.json > $O/train_n1.log 2>&1; echo "train n1 rc=$?"; tail -1 $O/train_n1.log | cut -c1-1800
