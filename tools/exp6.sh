#!/bin/bash
mkdir -p gpurun_out
KG_TC_ACC=1 KG_TC_MT=2 timeout 300 python tools/op_times.py 2>&1 | grep -E "^op  (58|60|62|64)|total" | tr '\n' ' ' ; echo " <- acc1 mt2"
KG_TC_STAGES=3 timeout 300 python tools/op_times.py 2>&1 | grep -E "^op  (58|60|62|64)|total" | tr '\n' ' ' ; echo " <- stages3"
timeout 300 python tools/op_times.py 2>&1 | grep -E "^op  (58|60|62|64)|total" | tr '\n' ' ' ; echo " <- default"
