#!/bin/bash
# ncu --set full of the memory-bound training kernels at the BENCHMARK batch (256 x 512x256 per GPU): achieved DRAM GB/s
set -u
O=gpurun_out
mkdir -p $O
OCTSEG_TRAIN_GRAPH=0 python tools/train_steps.py 256 2 > $O/plain_train256.log 2>&1 || exit 1
OCTSEG_TRAIN_GRAPH=0 ncu --set full --clock-control none -k regex:"bn_bwd|bn_finalize|pool_bwd|head_loss|wgrad_rows" -s 81 -c 81 -o $O/r2b_train_b256_ncu_full python tools/train_steps.py 256 2 > $O/ncu_b256.log 2>&1
python tools/ncu_summary.py $O/r2b_train_b256_ncu_full.ncu-rep $O/r2b_train_b256_ncu_full_summary.csv "launch (train step, batch 256, 512x256, bf16: BN forward / backward, pool backward, head + loss, row-walking weight gradients of one step)" && rm -f $O/r2b_train_b256_ncu_full.ncu-rep
ls -la $O | tail -5
