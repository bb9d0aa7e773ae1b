#!/bin/bash
mkdir -p gpurun_out
{
echo "=== fallbacks: OCTSEG_DGRAD_S2D=0 OCTSEG_WGRAD_ROWS=0"; OCTSEG_DGRAD_S2D=0 OCTSEG_WGRAD_ROWS=0 timeout 900 python -m pytest tests/test_gpu_backward_kernels.py tests/test_gpu_train.py -x -q -m gpu 2>&1 | tail -3
echo "=== default"; timeout 900 python -m pytest tests/test_gpu_backward_kernels.py tests/test_gpu_train.py -x -q -m gpu 2>&1 | tail -3
echo "=== train"; timeout 300 python tools/train_bench.py 256 20 | tail -1; timeout 300 python tools/train_bench.py 32 20 | tail -1
} > gpurun_out/r2b_run1.log 2>&1
cat gpurun_out/r2b_run1.log
