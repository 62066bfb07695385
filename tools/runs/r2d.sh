#!/bin/bash
# round 2, GPU call D: training row — backward kernels one by one, then the whole step
O=gpurun_out/r2d; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -s -x -k "wgrad or groupnorm_backward or layernorm_backward" > $O/ops.log 2>&1; echo "ops rc=$?"; tail -25 $O/ops.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -s -k "training_step_gradients or adam_and_ema" > $O/step.log 2>&1; echo "step rc=$?"; tail -40 $O/step.log | cut -c1-400
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -s -k "baseline_shape" > $O/b32.log 2>&1; echo "b32 rc=$?"; tail -15 $O/b32.log | cut -c1-400
