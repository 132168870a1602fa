#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2 3 7; do
  KG_SH_DBG=$d timeout 300 python tools/op_times.py 2>&1 | grep -E "^op  (59|61|63|65)" | tr '\n' ' ' > gpurun_out/exp3_dbg$d.log
  echo "dbg=$d: $(cat gpurun_out/exp3_dbg$d.log)"
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:tc_shift -s 8 -c 4 -f -o gpurun_out/exp3_shift python tools/op_times.py > gpurun_out/exp3_ncu.log 2>&1
ncu -i gpurun_out/exp3_shift.ncu-rep --page raw --csv > gpurun_out/exp3_shift_raw.csv 2>/dev/null
ncu -i gpurun_out/exp3_shift.ncu-rep --page source --csv --kernel-id :::1 > gpurun_out/exp3_shift_src.csv 2>/dev/null
ls -la gpurun_out/exp3*
