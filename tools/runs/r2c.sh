#!/bin/bash
# round 2, GPU call C (8 GPUs): BASELINE configs[2] (4096 patches, static split, full chains) and configs[3] (full frames -> .npy files)
O=gpurun_out/r2c; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus 8 --total-patches 4096 > $O/split4096_8gpu.json 2> $O/split.err; echo "split rc=$?"; cut -c1-900 $O/split4096_8gpu.json
timeout 300 $TR --master-port 29512 tools/bench_frame.py --frames 8 --out $O/frames8_8gpu.json > $O/frames8.log 2>&1; echo "frames8 rc=$?"; tail -2 $O/frames8.log | cut -c1-700
timeout 200 $TR --master-port 29513 tools/bench_frame.py --frames 1 --out $O/frames1_8gpu.json > $O/frames1.log 2>&1; echo "frames1 rc=$?"; tail -1 $O/frames1.log | cut -c1-700
nvidia-smi --query-gpu=index,name,clocks.sm,power.draw --format=csv > $O/smi.csv
