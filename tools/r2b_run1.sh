#!/bin/bash
mkdir -p gpurun_out
{
echo "=== tests"; timeout 900 python -m pytest tests/test_gpu_backward_kernels.py tests/test_gpu_train.py -x -q -m gpu 2>&1 | tail -3
echo "=== new"; timeout 300 python tools/predict_bench.py bf16
echo "=== old (df63f4a)"; OCTSEG_LIB=$PWD/tools/liboctseg_old.so timeout 300 python tools/predict_bench.py bf16
echo "=== new"; timeout 300 python tools/predict_bench.py bf16
echo "=== old (df63f4a)"; OCTSEG_LIB=$PWD/tools/liboctseg_old.so timeout 300 python tools/predict_bench.py bf16
} > gpurun_out/r2b_run1.log 2>&1
cat gpurun_out/r2b_run1.log
