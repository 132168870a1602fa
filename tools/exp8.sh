#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_net_gpu.py -q -k "shift or heads_l2 or forward" 2>&1 | tail -3
F="^op   (1|5):|^op  (53|56|59|61)|total"
timeout 300 python tools/op_times.py 2>&1 | grep -E "$F" | tr '\n' ' ' ; echo " <- default (2-pass wres)"
KG_TC_SHIFT_WRES=0 timeout 300 python tools/op_times.py 2>&1 | grep -E "$F" | tr '\n' ' ' ; echo " <- wres0 (2-pass streamed)"
KG_NO_2PASS=1 timeout 300 python tools/op_times.py 2>&1 | grep -E "$F" | tr '\n' ' ' ; echo " <- 3-pass"
KG_NO_2PASS=1 KG_TC_SHIFT_CONV=0 timeout 300 python tools/op_times.py 2>&1 | grep -E "$F" | tr '\n' ' ' ; echo " <- no shift conv"
