#!/bin/bash
# GPU check of the PDL train chain + row-walking weight gradient (round 2, second session)
mkdir -p gpurun_out
{
echo "=== hmma_bench"; timeout 60 tools/hmma_bench
echo "=== tests"; timeout 900 python -m pytest tests/test_gpu_backward_kernels.py tests/test_gpu_train.py -x -q 2>&1 | tail -15
for b in 256 32; do
  echo "=== train_bench $b default";           timeout 300 python tools/train_bench.py $b 10 2>&1 | tail -2
  echo "=== train_bench $b NO_PDL";            OCTSEG_NO_PDL=1 timeout 300 python tools/train_bench.py $b 10 2>&1 | tail -2
  echo "=== train_bench $b WGRAD_ROWS=0";      OCTSEG_WGRAD_ROWS=0 timeout 300 python tools/train_bench.py $b 10 2>&1 | tail -2
done
echo "=== phase profile 256 default"; OCTSEG_TRAIN_PROFILE=2 timeout 300 python tools/train_bench.py 256 2 2>&1 | grep -v "block .* \(conv_fwd\|bn_fwd\|dgrad\|bn_bwd\)" | tail -60
echo "=== phase profile 256 rows off"; OCTSEG_WGRAD_ROWS=0 OCTSEG_TRAIN_PROFILE=2 timeout 300 python tools/train_bench.py 256 2 2>&1 | grep -v "block .* \(conv_fwd\|bn_fwd\|dgrad\|bn_bwd\)" | tail -60
} > gpurun_out/r2b_run1.log 2>&1
tail -80 gpurun_out/r2b_run1.log
