#!/bin/bash
# round 2, GPU call AI: training step bench on the final tree (one GPU)
O=gpurun_out/r2ai; mkdir -p $O
timeout 600 python tools/bench_train.py --steps 10 --warmup 3 > $O/train_n1.json 2> $O/train.err; echo "rc=$?"; cut -c1-900 $O/train_n1.json; tail -2 $O/train.err
