#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:tc_shift -s 12 -c 2 -f -o gpurun_out/exp7_shift python tools/op_times.py > gpurun_out/exp7_ncu.log 2>&1
ncu -i gpurun_out/exp7_shift.ncu-rep --page raw --csv > gpurun_out/exp7_raw.csv 2>/dev/null
ncu -i gpurun_out/exp7_shift.ncu-rep --page source --csv --kernel-id :::1 > gpurun_out/exp7_src.csv 2>/dev/null
tail -3 gpurun_out/exp7_ncu.log
KG_TC_SHIFT_WRES=0 timeout 300 python tools/op_times.py 2>&1 | grep -E "^op   (1|5):|^op  (53|56)|total" | tr '\n' ' ' ; echo " <- wres0 (2-pass streamed)"
KG_NO_2PASS=1 KG_TC_SHIFT_WRES=0 timeout 300 python tools/op_times.py 2>&1 | grep -E "^op   (1|5):|^op  (53|56)|total" | tr '\n' ' ' ; echo " <- 3-pass streamed"
