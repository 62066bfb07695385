#!/bin/bash
# round 2, GPU call AQ: the child-process test of the two A/B switches
O=gpurun_out/r2aq; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "shared_memory_operand or plain_mma_form" > $O/tests.log 2>&1; echo "rc=$?"; tail -4 $O/tests.log | cut -c1-400
