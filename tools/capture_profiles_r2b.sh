#!/bin/bash
# Round-2 (second session) ncu evidence: train-step launch list and --set full of the kernels written this session.
# Every ncu run follows a plain run of the same command; reports are condensed on the box (gpurun_out/ <= 64 MiB).
set -u
O=gpurun_out
mkdir -p $O
sum() { python tools/ncu_summary.py $O/$1.ncu-rep $O/$1_summary.csv "$2" && rm -f $O/$1.ncu-rep; }
OCTSEG_TRAIN_GRAPH=0 python tools/train_steps.py 64 3 > $O/plain_train.log 2>&1 || exit 1
OCTSEG_TRAIN_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 190 -c 190 --csv --log-file $O/r2b_launches_train_step.csv python tools/train_steps.py 64 3 > $O/ncu_tr1.log 2>&1
OCTSEG_TRAIN_GRAPH=0 ncu --set full --import-source on --clock-control none -k regex:"wgrad_rows|head_loss|pool_bwd_add|wgrad_deep" -s 27 -c 27 -o $O/r2b_train_kernels_ncu_full python tools/train_steps.py 64 2 > $O/ncu_tr2.log 2>&1
sum r2b_train_kernels_ncu_full "launch (train step, per-GPU batch 64, 512x256: row-walking / deep weight gradients, head + loss, pool backward with BN reductions)"
# data-gradient convs of one step (conv_tc launches 22..43 of a step are the backward ones; capture all 43 of step 2)
OCTSEG_TRAIN_GRAPH=0 ncu --set full --clock-control none -k regex:conv_tc_kernel -s 43 -c 43 -o $O/r2b_train_convtc_ncu_full python tools/train_steps.py 64 2 > $O/ncu_tr3.log 2>&1
sum r2b_train_convtc_ncu_full "launch (train step, per-GPU batch 64, 512x256: the 43 conv_tc launches of one step -- 21 forward with statistics, 22 data gradients incl. the low-res up-conv ones)"
ls -la $O | tail -12
