#!/bin/bash
# round 2, GPU call AC: --set full source-level capture of ONE plain 64 -> 64 ws conv launch and ONE XF launch at the bench geometry
O=gpurun_out/r2ac; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:conv_gemm_kernel<\(int\)64, \(int\)4, \(bool\)1, \(bool\)0, \(bool\)1>' -s 2 -c 1 -o $O/conv_plain python tools/profile_step.py 64 2 > $O/ncu1.log 2>&1; echo "rc=$?"; tail -2 $O/ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:conv_gemm_kernel<\(int\)64, \(int\)4, \(bool\)1, \(bool\)1, \(bool\)1>' -s 2 -c 1 -o $O/conv_xf python tools/profile_step.py 64 2 > $O/ncu2.log 2>&1; echo "rc=$?"; tail -2 $O/ncu2.log
ls -la $O
