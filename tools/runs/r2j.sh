#!/bin/bash
# round 2, GPU call J (8 GPUs): BASELINE configs[4] — training step with the NCCL gradient all-reduce at 2 / 4 / 8 GPUs
O=gpurun_out/r2j; mkdir -p $O
for N in 2 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + N)) tools/bench_train.py --steps 10 --warmup 3 --out $O/train_n$N.json > $O/train_n$N.log 2>&1; echo "train N=$N rc=$?"; tail -1 $O/train_n$N.log | cut -c1-700
done
