#!/bin/bash
O=gpurun_out/r2h; mkdir -p $O
timeout 1500 python -m oracle.make_golden_stats_gpu $O/stats_1k_64_T1000.npz > $O/gen.log 2>&1; echo "gen rc=$?"; tail -12 $O/gen.log | cut -c1-400
