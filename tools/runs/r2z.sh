#!/bin/bash
# round 2, GPU call T: --set full source-level capture of the chain kernels after the elected-lane issue change
O=gpurun_out/r2z; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pixel_chain_kernel -s 4 -c 2 -o $O/chain_full python tools/profile_step.py 64 2 > $O/ncu.log 2>&1; echo "rc=$?"; tail -2 $O/ncu.log; ls -la $O
