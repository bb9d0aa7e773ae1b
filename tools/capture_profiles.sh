#!/bin/bash
# ncu evidence for profiles/ (run under gpurun; every ncu run follows a plain run of the same command line).
# Reports are condensed to CSV on the box and deleted: gpurun_out/ must stay under 64 MiB.
set -u
O=gpurun_out
mkdir -p $O
sum() { python tools/ncu_summary.py $O/$1.ncu-rep $O/$1_summary.csv "$2" && rm -f $O/$1.ncu-rep; }

# (1) wide net (BASELINE configs[3]): every conv_tc launch of one predict step, tensor-pipe counters
python tools/prof_predict.py bf16 2 wide > $O/plain_wide.log 2>&1 &&
ncu --set full --clock-control none -k regex:conv_tc_kernel -s 22 -c 22 -o $O/r2_wide_ncu_full python tools/prof_predict.py bf16 2 wide > $O/ncu_wide.log 2>&1
sum r2_wide_ncu_full "launch (wide U-Net predict, 8 x 1024x512, bf16: the 22 conv_tc launches of one step)"

# (2) train step (BASELINE configs[2] at per-GPU batch 64): launch list, then the memory-bound kernels + tcgen05 wgrad
OCTSEG_TRAIN_GRAPH=0 python tools/train_steps.py 64 3 > $O/plain_train.log 2>&1 &&
OCTSEG_TRAIN_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 190 -c 190 --csv --log-file $O/r2_launches_train_step.csv python tools/train_steps.py 64 3 > $O/ncu_tr1.log 2>&1
OCTSEG_TRAIN_GRAPH=0 python tools/train_steps.py 64 2 > $O/plain_train2.log 2>&1 &&
OCTSEG_TRAIN_GRAPH=0 ncu --set full --clock-control none -k regex:"bn_|head_loss|adam|wgrad_tc|pool_bwd" -s 66 -c 66 -o $O/r2_train_kernels_ncu_full python tools/train_steps.py 64 2 > $O/ncu_tr2.log 2>&1
sum r2_train_kernels_ncu_full "launch (train step, per-GPU batch 64, 512x256: BN / head-loss / pool-backward / Adam / tcgen05 wgrad launches of one step)"
ls -la $O | tail -12
