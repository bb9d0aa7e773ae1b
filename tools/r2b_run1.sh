#!/bin/bash
mkdir -p gpurun_out
{
echo "=== tests"; timeout 900 python -m pytest tests/test_gpu_backward_kernels.py tests/test_gpu_train.py -x -q 2>&1 | tail -5
for b in 256 32; do
  echo "=== train_bench $b"; timeout 300 python tools/train_bench.py $b 20 2>&1 | tail -1
done
OCTSEG_TRAIN_PROFILE=2 timeout 300 python tools/train_bench.py 256 2 2>&1 | sed 's/\[train detail\] //' > gpurun_out/prof_256.txt; grep "wgrad\|profile" gpurun_out/prof_256.txt | tail -23
} > gpurun_out/r2b_run1.log 2>&1
cat gpurun_out/r2b_run1.log
