#!/bin/bash
# per-block phase profile of the train step at per-GPU batch 32 and 256 (serialised on one stream, graph off)
mkdir -p gpurun_out
{
for b in 32 256; do
  echo "=== train_bench $b"; timeout 300 python tools/train_bench.py $b 20 2>&1 | tail -1
  echo "=== phase profile $b"; OCTSEG_TRAIN_PROFILE=2 timeout 300 python tools/train_bench.py $b 2 2>&1 | sed 's/\[train detail\] //' > gpurun_out/prof_$b.txt; tail -1 gpurun_out/prof_$b.txt
done
} > gpurun_out/r2b_run1.log 2>&1
cat gpurun_out/r2b_run1.log
