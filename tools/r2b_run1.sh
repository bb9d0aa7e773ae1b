#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
OCTSEG_TRAIN_GRAPH=0 python tools/train_steps.py 64 3 > $O/plain_train.log 2>&1 || exit 1
OCTSEG_TRAIN_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/r2b_launches_all.csv python tools/train_steps.py 64 3 > $O/ncu_tr1.log 2>&1
python tools/prof_predict.py bf16 3 > $O/plain_p.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/r2b_launches_predict_bf16_all.csv python tools/prof_predict.py bf16 3 > $O/ncu_p1.log 2>&1
python tools/prof_predict.py fp32 3 > $O/plain_p2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/r2b_launches_predict_fp32_all.csv python tools/prof_predict.py fp32 3 > $O/ncu_p2.log 2>&1
ls -la $O | tail -8
