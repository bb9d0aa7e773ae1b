#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 300 --warmup 10 --no-cpu-baseline --no-wide --no-fp32 2> gpurun_out/b8_err.log | tail -1 > gpurun_out/bench8.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench8.json').read())
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
print('train',d.get('train'))
print('cfg5',d.get('cfg5'))
PY
tail -3 gpurun_out/b8_err.log
