#!/bin/bash
O=gpurun_out/r2l; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_frames.py -m gpu -q -s -k "other_geometries or compose_noisy" > $O/new_tests.log 2>&1; echo "new tests rc=$?"; tail -8 $O/new_tests.log | cut -c1-400
# the dominant kernel, --set full: six conv_gemm launches of the second reverse step (graph-free eager launches)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 80 -c 6 -o $O/conv_full python tools/profile_step.py 64 2 > $O/ncu2.log 2>&1; echo "set full rc=$?"; tail -3 $O/ncu2.log; ls -la $O
