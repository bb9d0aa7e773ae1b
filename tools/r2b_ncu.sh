#!/bin/bash
# ncu --set full of the head launch (<4,1>) and the stem (<0,3>) of one bf16 predict step, with source lines and stall reasons
mkdir -p gpurun_out
O=gpurun_out
python tools/prof_predict.py bf16 2 > $O/plain.log 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:conv_tc_kernel -s 22 -c 22 -o $O/headstem python tools/prof_predict.py bf16 2 > $O/ncu.log 2>&1
ncu -i $O/headstem.ncu-rep --page raw --csv > $O/headstem_raw.csv 2>/dev/null
ncu -i $O/headstem.ncu-rep --page source --print-source cuda,sass --csv > $O/headstem_source.csv 2>/dev/null
ncu -i $O/headstem.ncu-rep --page details > $O/headstem_details.txt 2>/dev/null
rm -f $O/headstem.ncu-rep
ls -la $O | tail
