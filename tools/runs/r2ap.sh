#!/bin/bash
# round 2, GPU call AP: two-GPU run of the default bench on the final tree (the driver's launch line)
O=gpurun_out/r2ap; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "rc=$?"; tail -1 $O/bench_n2.json | cut -c1-500; tail -2 $O/bench_n2.err | cut -c1-300
