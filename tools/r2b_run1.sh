#!/bin/bash
mkdir -p gpurun_out
{
for v in 0 1 0 1; do
  echo "=== OCTSEG_ONE_D2H_STREAM=$v"; OCTSEG_ONE_D2H_STREAM=$v timeout 300 python bench.py --steps 300 --no-train --no-wide --no-cfg5 --no-fp32 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',round(d['value']),'e2e',round(d['e2e']['value']),'sync',round(d['e2e']['sync_call']['value']),'probs',round(d['e2e_probs']['value']))"
done
echo "=== api tests"; timeout 600 python -m pytest tests/test_gpu_api.py -x -q -m gpu 2>&1 | tail -3
} > gpurun_out/r2b_run1.log 2>&1
cat gpurun_out/r2b_run1.log
