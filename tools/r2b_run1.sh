#!/bin/bash
mkdir -p gpurun_out
{
echo "=== tests"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_gpu_train.py -x -q -m gpu 2>&1 | tail -3
echo "=== predict"; timeout 300 python tools/predict_bench.py fp32 bf16
echo "=== train"; timeout 300 python tools/train_bench.py 256 20 | tail -1
OCTSEG_TRAIN_PROFILE=2 timeout 300 python tools/train_bench.py 256 2 2>&1 | sed 's/\[train detail\] //' | grep "block 0 conv_fwd" | tail -1
} > gpurun_out/r2b_run1.log 2>&1
cat gpurun_out/r2b_run1.log
