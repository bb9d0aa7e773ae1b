#!/bin/bash
mkdir -p gpurun_out
{
nvidia-smi -L
echo "=== 2-GPU tests"; timeout 600 python -m pytest tests/test_gpu_baseline_shapes.py -x -q -m gpu -k "two" 2>&1 | tail -5
echo "=== bench --gpus 2 (train + predict, short)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 --no-cpu-baseline --no-wide --no-cfg5 --no-fp32 2> gpurun_out/b2_err.log | tail -1 > gpurun_out/bench2.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench2.json').read())
print('value',d['value'],'e2e',d['e2e']['value'],'train',d.get('train',{}).get('value'),d.get('train',{}).get('ms_per_step'))
PY
tail -3 gpurun_out/b2_err.log
} > gpurun_out/r2b_run2gpu.log 2>&1
cat gpurun_out/r2b_run2gpu.log
