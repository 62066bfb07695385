#!/bin/bash
O=gpurun_out/r2e; mkdir -p $O
timeout 600 python tools/bench_train.py --steps 10 --warmup 3 --out $O/train_n1.json > $O/train_n1.log 2>&1; echo "train n1 rc=$?"; tail -3 $O/train_n1.log | cut -c1-2500
