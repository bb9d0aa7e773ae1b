#!/bin/bash
mkdir -p gpurun_out
{
for rep in 1 2; do
for v in 0 1; do
  echo "=== OCTSEG_PDL_STREAM=$v"; OCTSEG_PDL_STREAM=$v timeout 300 python tools/train_bench.py 32 30 | tail -1; OCTSEG_PDL_STREAM=$v timeout 300 python tools/train_bench.py 256 20 | tail -1
done
done
echo "=== tests with OCTSEG_PDL_STREAM=1"; OCTSEG_PDL_STREAM=1 timeout 900 python -m pytest tests/test_gpu_train.py -x -q -m gpu 2>&1 | tail -3
} > gpurun_out/r2b_run1.log 2>&1
cat gpurun_out/r2b_run1.log
