#!/bin/bash
# quick check of a build: every -m gpu test, smoke, one train-step timing
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -2
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 200 python tools/train_bench.py 256 20 | tail -1
} > gpurun_out/r2b_quick.log 2>&1
cat gpurun_out/r2b_quick.log
