#!/bin/bash
# round 2, GPU call I: final evidence for the sampling path — launch list with DRAM bytes (plan-signed), --set full of the top kernel
O=gpurun_out/r2i; mkdir -p $O
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active
timeout 900 ncu --metrics $M --clock-control none --csv --log-file $O/launches_step_mb64.csv python tools/profile_step.py 64 2 > $O/ncu1.log 2>&1; echo "launch list rc=$?"; tail -1 $O/ncu1.log
timeout 300 python bench.py --no-e2e --no-cpu --steps 10 --warmup 3 --dump-layers $O/layers.json > $O/bench_short.json 2> $O/bench.err; echo "bench rc=$?"
python tools/ncu_launch_summary.py $O/launches_step_mb64.csv $O/layers.json $O/ncu_step_summary.json
# the dominant kernel, --set full, three launches of the 64 -> 64 ws variant and three of the N = 128 streamed variant
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 80 -c 6 -o $O/conv_full python tools/profile_step.py 64 2 > $O/ncu2.log 2>&1; echo "set full rc=$?"; ls -la $O
