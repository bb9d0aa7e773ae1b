#!/bin/bash
mkdir -p gpurun_out
{
echo "=== 12 epilogue warps"; timeout 300 python tools/predict_bench.py bf16 fp32
echo "=== 16 epilogue warps"; OCTSEG_LIB=$PWD/tools/liboctseg_epi16.so timeout 300 python tools/predict_bench.py bf16 fp32
echo "=== 12 again"; timeout 300 python tools/predict_bench.py bf16 fp32
} > gpurun_out/r2b_run1.log 2>&1
cat gpurun_out/r2b_run1.log
