#!/bin/bash
for lv in 5 3 2; do
echo "== KG_SEG_1PASS_LEVELS=$lv"
KG_SEG_1PASS_LEVELS=$lv timeout 300 python -m pytest tests/test_net_gpu.py -q -k "seg" 2>&1 | grep -E "assert|passed|failed|worst" | head -5
done
KG_SEG_1PASS_LEVELS=5 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['stages']['forward_seg'])"
