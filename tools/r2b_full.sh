#!/bin/bash
# full GPU validation: every -m gpu test, smoke(), the default bench line and the reference arm
mkdir -p gpurun_out
{
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -8
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "=== bench"; timeout 900 python bench.py > gpurun_out/bench_line.json 2> gpurun_out/bench_err.log; echo "rc=$?"; tail -3 gpurun_out/bench_err.log
echo "=== bench reference"; timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_line.json 2>> gpurun_out/bench_err.log; echo "rc=$?"
} > gpurun_out/r2b_full.log 2>&1
cat gpurun_out/r2b_full.log
