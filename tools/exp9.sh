#!/bin/bash
F="^op   (1|5):|^op  (53|56|59|61|63|65)|total"
timeout 300 python -m pytest tests/test_net_gpu.py -q -k "shift or heads_l2 or forward" 2>&1 | tail -2
for pf in 0 1 2 4; do
KG_TC_SHIFT_PF=$pf timeout 300 python tools/op_times.py 2>&1 | grep -E "$F" | tr '\n' ' ' ; echo " <- pf=$pf"
done
