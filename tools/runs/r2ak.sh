#!/bin/bash
# round 2, GPU call AK: --set full source-level capture of the init_conv launch (kDirect, Toeplitz operand) at the bench geometry
O=gpurun_out/r2ak; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:conv_gemm_kernel<\(int\)64, \(int\)0, \(bool\)1, \(bool\)0, \(bool\)0>' -s 1 -c 1 -o $O/init_conv python tools/profile_step.py 64 2 > $O/ncu1.log 2>&1; echo "rc=$?"; tail -2 $O/ncu1.log; ls -la $O
