#!/bin/bash
export CUDA_LAUNCH_BLOCKING=1
for k in test_cross_entropy test_silog test_upsample_bilinear_backward test_batchnorm_training test_conv3x3_weight test_training_forward test_training_step; do
  echo "=== $k"
  timeout 200 python -m pytest tests/test_gpu_train_tail.py -m gpu -q -s -x -k $k 2>&1 | grep -v "^$" | grep "^E  \|passed\|failed\|Error\|error\|worst\|assert" | head -14
done
