#!/bin/bash
O=gpurun_out/r2k; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_stats.py -m gpu -q -s -k "full_chain_64" > $O/stats64.log 2>&1; echo "stats64 rc=$?"; tail -12 $O/stats64.log | cut -c1-600
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -x > $O/train.log 2>&1; echo "train rc=$?"; tail -3 $O/train.log
timeout 400 python tools/bench_train.py --steps 10 --warmup 3 --out $O/train_n1.json > $O/train_n1.log 2>&1; echo "bench_train rc=$?"; tail -1 $O/train_n1.log | cut -c1-1600
bash tools/runs/r2i.sh
